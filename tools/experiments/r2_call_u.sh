# round 2, GPU call U (1 GPU): the whole GPU test suite on the final tree + the --set full capture of the launch the tensor-pipe target is
# judged on (res5b branch2b, dilated 3x3, now in CTA-pair form)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1
tail -4 gpurun_out/r2u_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -3 gpurun_out/r2u_smoke.log
export DC_CUDA_GRAPH=0
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip 150 --launch-count 1 -f -o gpurun_out/r2u_prof_res5b_2b \
   python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2u_ncu_res5b_2b.log 2>&1
ls -la gpurun_out | grep r2u
