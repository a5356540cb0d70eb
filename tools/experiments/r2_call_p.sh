# round 2, GPU call P (1 GPU): pairs for the 3x3 convs as the default -- tests, then re-measure the tile choices that were made while
# MMA issue was the bottleneck (256-channel tiles, lean epilogue), batch-1 latency, bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_tests.log 2>&1
tail -4 gpurun_out/r2p_tests.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2p_sweep.jsonl \
  --config "base:" --config "bn256_everywhere:DC_CONV_BN256=2" --config "bn256_off:DC_CONV_BN256=0" --config "lean_off:DC_LEAN_EPILOGUE=0" \
  --config "pair3x3_off:DC_CONV_PAIR_3X3=0" --config "no_weight_hint:DC_WEIGHTS_EVICT_LAST=0" --config "base_again:" > gpurun_out/r2p_sweep.log 2>&1
tail -3 gpurun_out/r2p_sweep.log | cut -c1-200
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2p_sweep.jsonl')]
keys=sorted(rows[0]['stage_ms'])
print('%-10s'%'stage', *['%10s'%r['config'][:10] for r in rows])
for k in keys: print('%-10s'%k, *['%10.3f'%r['stage_ms'].get(k,0) for r in rows])
print('%-10s'%'step', *['%10.3f'%r['ms_per_step'] for r in rows])
print('maxdiff', *[max(r['max_abs_diff_vs_first'].values()) for r in rows])
PY
timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2p_cfg1.json 2> gpurun_out/r2p_cfg1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2p_cfg1.json').read()); print('cfg1', d['ms_per_step'], d['e2e']['value'])"
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2p_steps_16x720p.json > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench_n1.json').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['without_next_pred'], d['clocks'], d['roofline']['frac'], d['roofline']['issued_frac'], d['cpu_baseline'])"
