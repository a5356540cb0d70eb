# GPU call H: split-K exchange through the L2-resident scratch + remote mbarrier arrive: correctness, latency, cfg1.
set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" ) > gpurun_out/h_tests_splitk.log 2>&1
tail -5 gpurun_out/h_tests_splitk.log
for cfg in "1 36" "4 36" "4 16" "2 16"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SPLIT_K_MIN_STEPS=$2 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency
done > gpurun_out/h_lat.txt
cat gpurun_out/h_lat.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/h_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/h_tests.log | tail -5
for cfg in "1 36" "4 36" "4 16"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_SPLIT_K_MIN_STEPS=$2 timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/h_bench_cfg1_sk$1_min$2.json 2> gpurun_out/h_bench_cfg1_sk$1_min$2.err
  cut -c1-200 gpurun_out/h_bench_cfg1_sk$1_min$2.json; tail -3 gpurun_out/h_bench_cfg1_sk$1_min$2.err
done
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" > gpurun_out/h_memcheck.log 2>&1
tail -4 gpurun_out/h_memcheck.log
