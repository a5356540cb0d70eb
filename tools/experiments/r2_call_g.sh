# round 2, GPU call G (1 GPU): 256-channel-tile pair kernel for the 1x1 reduce convs -- parity, A/B timing, ncu
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1
tail -5 gpurun_out/r2g_tests.log
DC_CONV_BN256=0 timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2g_sweep.jsonl --config "bn128:" --config "bn128_again:" > gpurun_out/r2g_sweep_a.log 2>&1
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2g_sweep.jsonl --config "bn256:" --config "bn256_again:" > gpurun_out/r2g_sweep_b.log 2>&1
DC_CONV_BN256=0 timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2g_sweep.jsonl --config "bn128_third:" > gpurun_out/r2g_sweep_c.log 2>&1
cat gpurun_out/r2g_sweep.jsonl | cut -c1-400
export DC_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip 58 --launch-count 1 -f -o gpurun_out/r2g_prof_res4b7_2a_bn256 \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2g_ncu_res4b7_2a.log 2>&1
ls -la gpurun_out | tail -8
