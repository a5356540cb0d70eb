# round 2, GPU call H (1 GPU): what bounds the 1x1 convs (operand-skip microbenchmark), kernel tests with the wave-aware BN=256 choice
set -x
mkdir -p gpurun_out
timeout 900 python tools/conv_microbench.py --set bound > gpurun_out/r2h_microbench_bound.txt 2>&1
cat gpurun_out/r2h_microbench_bound.txt
DC_CONV_BN256=2 timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/r2h_tests_kernels_bn256.log 2>&1
tail -3 gpurun_out/r2h_tests_kernels_bn256.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1
tail -3 gpurun_out/r2h_tests.log
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2h_sweep.jsonl --config "final:" --config "final_again:" > gpurun_out/r2h_sweep.log 2>&1
cut -c1-300 gpurun_out/r2h_sweep.jsonl
