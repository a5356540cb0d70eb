# GPU call J: split-K through L2 scratch with batched slot loads.
set -x
mkdir -p gpurun_out
for cfg in "1 36" "4 16" "2 16"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SPLIT_K_MIN_STEPS=$2 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency
done > gpurun_out/j_lat.txt
cat gpurun_out/j_lat.txt
( time timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_net_gpu.py -m gpu -x -q -k "split_k or batch_independence or resnet152" ) > gpurun_out/j_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/j_tests.log | tail -3
for cfg in "1 36" "4 36" "4 16" "4 32"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SPLIT_K_MIN_STEPS=$2 timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/j_bench_cfg1_sk$1_min$2.json 2> gpurun_out/j_bench_cfg1_sk$1_min$2.err
  cut -c1-200 gpurun_out/j_bench_cfg1_sk$1_min$2.json; tail -3 gpurun_out/j_bench_cfg1_sk$1_min$2.err
done
