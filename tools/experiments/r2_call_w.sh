# round 2, GPU call W (1 GPU): compute-sanitizer memcheck over the kernel tests that exercise the round's new kernel forms (TALL mode, CTA
# pairs for the 3x3 convs, 256-channel tiles under force, the stem in TALL mode)
set -x
mkdir -p gpurun_out
timeout 330 compute-sanitizer --tool memcheck --print-limit 20 --log-file gpurun_out/r2w_memcheck.log \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "case17 or case19 or case20 or case21 or case22 or test_conv1_stem_tensor_core or (neutral and case4)" > gpurun_out/r2w_tests.log 2>&1
tail -3 gpurun_out/r2w_tests.log
tail -5 gpurun_out/r2w_memcheck.log
