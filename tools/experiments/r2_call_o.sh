# round 2, GPU call O (1 GPU): tests after the weight-cache fix (skip_outputs after a full forward), then A/B on the 16x720p step:
# HBM->L2 prefetch distance for the stride-1 1x1 convs, CTA pairs for the 3x3 convs now that MMA issue is cheap
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1
tail -4 gpurun_out/r2o_tests.log
grep -h "parity\] survey" gpurun_out/r2o_tests.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2o_sweep.jsonl \
  --config "base:" --config "pf4:DC_L2_PREFETCH=4" --config "pf8:DC_L2_PREFETCH=8" --config "pf16:DC_L2_PREFETCH=16" --config "pf32:DC_L2_PREFETCH=32" \
  --config "pair3x3:DC_CONV_PAIR_3X3=1" --config "pair_all:DC_CONV_PAIR_ALL=1" --config "base_again:" > gpurun_out/r2o_sweep.log 2>&1
tail -3 gpurun_out/r2o_sweep.log
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/r2o_sweep.jsonl')]
keys=sorted(rows[0]['stage_ms'])
print('%-10s'%'stage', *['%10s'%r['config'][:10] for r in rows])
for k in keys: print('%-10s'%k, *['%10.3f'%r['stage_ms'].get(k,0) for r in rows])
print('%-10s'%'step', *['%10.3f'%r['ms_per_step'] for r in rows])
print('maxdiff', *[max(r['max_abs_diff_vs_first'].values()) for r in rows])
PY
