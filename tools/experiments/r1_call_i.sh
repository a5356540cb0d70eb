# GPU call I: split-K through L2 scratch -- which fence makes it slow?
set -x
mkdir -p gpurun_out
for f in 0 1 2; do
  DC_SK_FENCE=$f DC_SPLIT_K=4 DC_SPLIT_K_MIN_STEPS=16 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency | sed "s/\$/ sk_fence=$f/"
done > gpurun_out/i_lat.txt
cat gpurun_out/i_lat.txt
( time timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_net_gpu.py -m gpu -x -q -k "split_k or batch_independence or resnet152" ) > gpurun_out/i_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/i_tests.log | tail -3
for cfg in "4 36" "4 16"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SPLIT_K_MIN_STEPS=$2 timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/i_bench_cfg1_sk$1_min$2.json 2> gpurun_out/i_bench_cfg1_sk$1_min$2.err
  cut -c1-200 gpurun_out/i_bench_cfg1_sk$1_min$2.json; tail -3 gpurun_out/i_bench_cfg1_sk$1_min$2.err
done
