# GPU call C: where does the split-K overhead come from?  Latency microbench (graph-captured dependent chains).
set -x
mkdir -p gpurun_out
for cfg in "1 0" "2 0" "4 0" "4 1" "4 3" "2 1" "2 3"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SK_DEBUG=$2 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency
done > gpurun_out/c_lat.txt
cat gpurun_out/c_lat.txt
DC_SMALL_GRID_BN64=0 DC_SPLIT_K=1 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency > gpurun_out/c_lat_bn128.txt
cat gpurun_out/c_lat_bn128.txt
