# round 2, GPU call D (1 GPU): tests on the tree with L2 hints / serpentine plumbing (both off by default), hint sweep
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1
tail -5 gpurun_out/r2d_tests.log
timeout 900 python tools/chunk_sweep.py --out gpurun_out/r2d_sweep.jsonl \
  --config "default:" \
  --config "serp:DC_SERPENTINE=1" \
  --config "w_last:DC_L2_HINTS=2" \
  --config "out_last:DC_L2_HINTS=4" \
  --config "out_last_w:DC_L2_HINTS=6" \
  --config "first:DC_L2_HINTS=1" \
  --config "all:DC_L2_HINTS=7" \
  --config "all_serp:DC_L2_HINTS=7;DC_SERPENTINE=1" \
  --config "all_32:DC_L2_HINTS=7;DC_L2_HINT_MB=32" \
  --config "out_last_serp:DC_L2_HINTS=4;DC_SERPENTINE=1" \
  --config "all_chunk8:DC_L2_HINTS=7;DC_CHUNK_PLAN=0,4,8,4" \
  --config "default_again:" \
  > gpurun_out/r2d_sweep.log 2>&1
tail -2 gpurun_out/r2d_sweep.log
