# GPU call D: split-K with bulk DSMEM exchange: correctness, latency microbench, BASELINE configs[1].
set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" ) > gpurun_out/d_tests_splitk.log 2>&1
tail -5 gpurun_out/d_tests_splitk.log
for v in 1 2 4; do
  DC_SPLIT_K=$v timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency
done > gpurun_out/d_lat.txt
cat gpurun_out/d_lat.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/d_tests.log 2>&1
tail -5 gpurun_out/d_tests.log
for v in 1 4; do
  DC_SPLIT_K=$v timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline --step-report gpurun_out/d_steps_cfg1_sk$v.json > gpurun_out/d_bench_cfg1_sk$v.json 2> gpurun_out/d_bench_cfg1_sk$v.err
  cut -c1-330 gpurun_out/d_bench_cfg1_sk$v.json; tail -3 gpurun_out/d_bench_cfg1_sk$v.err
done
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" > gpurun_out/d_memcheck.log 2>&1
tail -4 gpurun_out/d_memcheck.log
