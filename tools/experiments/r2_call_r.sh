# round 2, GPU call R (1 GPU): TALL mode for the 64 -> 64 channel convs (res2 branch2b, stem) -- tests, A/B on the 16x720p step
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2r_tests_kernels.log 2>&1
tail -4 gpurun_out/r2r_tests_kernels.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_kernels_gpu.py > gpurun_out/r2r_tests.log 2>&1
tail -4 gpurun_out/r2r_tests.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2r_sweep.jsonl \
  --config "tall:" --config "tall_off:DC_CONV_TALL=0" --config "tall_again:" > gpurun_out/r2r_sweep.log 2>&1
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
timeout 300 python tools/conv_microbench.py --set res2 > gpurun_out/r2r_micro_res2_tall.txt 2>&1
DC_CONV_TALL=0 timeout 300 python tools/conv_microbench.py --set res2 > gpurun_out/r2r_micro_res2_plain.txt 2>&1
cat gpurun_out/r2r_micro_res2_tall.txt gpurun_out/r2r_micro_res2_plain.txt
