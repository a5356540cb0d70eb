# round 2, GPU call Q (1 GPU): 256-channel tiles off by default + new kernel tests; A/B: a pinned fraction of every residual-block
# output in L2 (fractional evict_last on the lean epilogue's TMA stores)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1
tail -4 gpurun_out/r2q_tests.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2q_sweep.jsonl \
  --config "base:" --config "keep32:DC_OUT_KEEP_MB=32" --config "keep56:DC_OUT_KEEP_MB=56" --config "keep80:DC_OUT_KEEP_MB=80" --config "keep104:DC_OUT_KEEP_MB=104" \
  --config "base_again:" > gpurun_out/r2q_sweep.log 2>&1
tail -3 gpurun_out/r2q_sweep.log | cut -c1-200
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2q_sweep.jsonl')]
keys=sorted(rows[0]['stage_ms'])
print('%-10s'%'stage', *['%10s'%r['config'][:10] for r in rows])
for k in keys: print('%-10s'%k, *['%10.3f'%r['stage_ms'].get(k,0) for r in rows])
print('%-10s'%'step', *['%10.3f'%r['ms_per_step'] for r in rows])
print('maxdiff', *[max(r['max_abs_diff_vs_first'].values()) for r in rows])
PY
