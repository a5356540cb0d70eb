# round 2, GPU call R (1 GPU): TALL mode for the 64 -> 64 channel convs (res2 branch2b, stem) -- tests, A/B on the 16x720p step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2r_tests_kernels.log 2>&1
tail -4 gpurun_out/r2r_tests_kernels.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_kernels_gpu.py > gpurun_out/r2r_tests.log 2>&1
tail -4 gpurun_out/r2r_tests.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2r_sweep.jsonl \
  --config "tall:" --config "tall_off:DC_CONV_TALL=0" --config "tall_again:" --config "tall_off_again:DC_CONV_TALL=0" > gpurun_out/r2r_sweep.log 2>&1
tail -3 gpurun_out/r2r_sweep.log | cut -c1-200
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2r_sweep.jsonl')]
keys=sorted(rows[0]['stage_ms'])
print('%-10s'%'stage', *['%10s'%r['config'][:10] for r in rows])
for k in keys: print('%-10s'%k, *['%10.3f'%r['stage_ms'].get(k,0) for r in rows])
print('%-10s'%'step', *['%10.3f'%r['ms_per_step'] for r in rows])
print('maxdiff', *[max(r['max_abs_diff_vs_first'].values()) for r in rows])
PY



timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2r_cfg1.json 2> gpurun_out/r2r_cfg1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2r_cfg1.json').read()); print('cfg1', d['ms_per_step'], d['e2e']['value'])"
