# round 2, GPU call W2 (1 GPU): compute-sanitizer memcheck over ALL kernel tests and the small whole-net tests of the final tree
set -x
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 --log-file gpurun_out/r2w2_memcheck_kernels.log \
  python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2w2_tests_kernels.log 2>&1
tail -3 gpurun_out/r2w2_tests_kernels.log; tail -3 gpurun_out/r2w2_memcheck_kernels.log
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 --log-file gpurun_out/r2w2_memcheck_net.log \
  python -m pytest tests/test_net_gpu.py tests/test_api_gpu.py -x -q -k "tiny or skipped or ragged or trimmed" > gpurun_out/r2w2_tests_net.log 2>&1
tail -3 gpurun_out/r2w2_tests_net.log; tail -3 gpurun_out/r2w2_memcheck_net.log
