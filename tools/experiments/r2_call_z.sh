# round 2, GPU call Z (1 GPU): smoke() + the TALL / pair kernel tests on the committed final tree (after the comment-only rebuild)
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -3 gpurun_out/r2z_smoke.log
timeout 60 python -m pytest tests/test_kernels_gpu.py -x -q -k "neutral or case20 or case17" > gpurun_out/r2z_tests.log 2>&1; tail -2 gpurun_out/r2z_tests.log
