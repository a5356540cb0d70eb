# round 2, GPU call B: double-buffered lean epilogue (2c convs) -- parity, sweep, bench line, ncu record
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1
tail -5 gpurun_out/r2b_tests.log
timeout 900 python tools/chunk_sweep.py --out gpurun_out/r2b_sweep.jsonl \
  --config "default:" \
  --config "noinplace:DC_INPLACE_RESIDUAL=0" \
  --config "c0_2_4_2:DC_CHUNK_PLAN=0,2,4,2" \
  --config "c0_4_8_4:DC_CHUNK_PLAN=0,4,8,4" \
  --config "c0_0_4_0:DC_CHUNK_PLAN=0,0,4,0" \
  --config "c0_0_8_0:DC_CHUNK_PLAN=0,0,8,0" \
  --config "default_again:" \
  > gpurun_out/r2b_sweep.log 2>&1
tail -2 gpurun_out/r2b_sweep.log
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2b_steps_16x720p.json > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
cat gpurun_out/r2b_bench_n1.json
export DC_CUDA_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2b_launches_16x720p.csv python tools/profile_forward.py --warm 2 --iters 1 --profiler-range --schedule-out gpurun_out/r2b_schedule.txt > gpurun_out/r2b_ncu_l.log 2>&1
for spec in res4b7_2a:58 res4b7_2c:60 res2b_2c:7; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip $idx --launch-count 1 -f -o gpurun_out/r2b_prof_$name \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2b_ncu_$name.log 2>&1
done
for spec in head_finish:2 maxpool_split:0 stem_s2d:0; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$name --launch-skip $idx --launch-count 1 -f -o gpurun_out/r2b_prof_$name \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2b_ncu_$name.log 2>&1
done
ls -la gpurun_out | tail -30
