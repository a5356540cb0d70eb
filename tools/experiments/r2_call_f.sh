# round 2, GPU call F (1 GPU): (1) parity cost of merging the cross terms into the main accumulator on the 1x1 reduce convs
# (what BN = 256 tiles with double-buffered accumulators would need); (2) sweep re-check of the L2 hint default
set -x
mkdir -p gpurun_out
DC_PARITY_JSON=$PWD/gpurun_out/r2f_parity_default.json timeout 900 python -m pytest tests/test_net_gpu.py -q -k "resnet152_matches" > gpurun_out/r2f_parity_default.log 2>&1
DC_MERGE_ACC_2A=1 DC_PARITY_JSON=$PWD/gpurun_out/r2f_parity_merge2a.json timeout 900 python -m pytest tests/test_net_gpu.py -q -k "resnet152_matches" > gpurun_out/r2f_parity_merge2a.log 2>&1
tail -3 gpurun_out/r2f_parity_default.log gpurun_out/r2f_parity_merge2a.log
timeout 900 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2f_sweep.jsonl \
  --config "default_w_last:" \
  --config "none:DC_L2_HINTS=0" \
  --config "all:DC_L2_HINTS=7" \
  --config "merge2a:DC_MERGE_ACC_2A=1" \
  --config "none_again:DC_L2_HINTS=0" \
  --config "default_again:" \
  > gpurun_out/r2f_sweep.log 2>&1
tail -2 gpurun_out/r2f_sweep.log
