# round 2, GPU call C (1 GPU): full GPU tests on the final kernels, bench line, pyramid at N=1, ncu of the restored lean kernel
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1
tail -5 gpurun_out/r2c_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2c_steps_16x720p.json > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
cat gpurun_out/r2c_bench_n1.json
timeout 600 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_cfg1.json 2> gpurun_out/r2c_bench_cfg1.err
timeout 600 python bench.py --workload batch128_512 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_cfg3_n1.json 2> gpurun_out/r2c_bench_cfg3_n1.err
timeout 900 python tools/pyramid_bench.py --steps 5 > gpurun_out/r2c_pyramid_n1.json 2> gpurun_out/r2c_pyramid_n1.err
cat gpurun_out/r2c_pyramid_n1.json
export DC_CUDA_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2c_launches_16x720p.csv python tools/profile_forward.py --warm 2 --iters 1 --profiler-range --schedule-out gpurun_out/r2c_schedule.txt > gpurun_out/r2c_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip 60 --launch-count 1 -f -o gpurun_out/r2c_prof_res4b7_2c \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2c_ncu_res4b7_2c.log 2>&1
ls -la gpurun_out | tail -12
