# GPU call B: split-K correctness (kernel + net tests), then BASELINE configs[1] with DC_SPLIT_K = 1 / 2 / 4.
set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" ) > gpurun_out/b_tests_splitk.log 2>&1
tail -15 gpurun_out/b_tests_splitk.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/b_tests.log 2>&1
tail -15 gpurun_out/b_tests.log
for v in 1 2 4; do
  DC_SPLIT_K=$v timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline --step-report gpurun_out/b_steps_cfg1_sk$v.json > gpurun_out/b_bench_cfg1_sk$v.json 2> gpurun_out/b_bench_cfg1_sk$v.err
  cut -c1-420 gpurun_out/b_bench_cfg1_sk$v.json; tail -3 gpurun_out/b_bench_cfg1_sk$v.err
done
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "split_k" > gpurun_out/b_memcheck.log 2>&1
tail -8 gpurun_out/b_memcheck.log
