# round 2, GPU call A: parity of the chunked L2-resident schedule, the schedule sweep, DRAM traffic per forward under ncu
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
tail -5 gpurun_out/r2a_tests.log
timeout 900 python tools/chunk_sweep.py --out gpurun_out/r2a_sweep.jsonl \
  --config "r1like:DC_CHUNK_PLAN=0,0,0,0;DC_INPLACE_RESIDUAL=0" \
  --config "inplace:DC_CHUNK_PLAN=0,0,0,0" \
  --config "c0_1_3_1:DC_CHUNK_PLAN=0,1,3,1" \
  --config "c0_2_4_2:DC_CHUNK_PLAN=0,2,4,2" \
  --config "c0_2_2_2:DC_CHUNK_PLAN=0,2,2,2" \
  --config "c0_4_8_4:DC_CHUNK_PLAN=0,4,8,4" \
  --config "c0_0_2_0:DC_CHUNK_PLAN=0,0,2,0" \
  --config "c0_0_3_0:DC_CHUNK_PLAN=0,0,3,0" \
  --config "c0_0_4_0:DC_CHUNK_PLAN=0,0,4,0" \
  --config "c0_0_6_0:DC_CHUNK_PLAN=0,0,6,0" \
  --config "c0_0_8_0:DC_CHUNK_PLAN=0,0,8,0" \
  --config "c0_1_0_0:DC_CHUNK_PLAN=0,1,0,0" \
  --config "c0_2_0_0:DC_CHUNK_PLAN=0,2,0,0" \
  --config "c0_4_0_0:DC_CHUNK_PLAN=0,4,0,0" \
  --config "c0_0_0_1:DC_CHUNK_PLAN=0,0,0,1" \
  --config "c0_0_0_2:DC_CHUNK_PLAN=0,0,0,2" \
  --config "c0_0_0_4:DC_CHUNK_PLAN=0,0,0,4" \
  --config "c0_2_4_2_noinplace:DC_CHUNK_PLAN=0,2,4,2;DC_INPLACE_RESIDUAL=0" \
  --config "r1like_again:DC_CHUNK_PLAN=0,0,0,0;DC_INPLACE_RESIDUAL=0" \
  > gpurun_out/r2a_sweep.log 2>&1
tail -3 gpurun_out/r2a_sweep.log
export DC_CUDA_GRAPH=0
for cfg in "r1like 0,0,0,0 0" "c0_2_4_2 0,2,4,2 1" "c0_1_3_1 0,1,3,1 1"; do
  set -- $cfg
  DC_CHUNK_PLAN=$2 DC_INPLACE_RESIDUAL=$3 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
     --cache-control none --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2a_dram_$1.csv \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range --schedule-out gpurun_out/r2a_schedule_$1.txt > gpurun_out/r2a_ncu_$1.log 2>&1
  tail -2 gpurun_out/r2a_ncu_$1.log
done
ls -la gpurun_out | tail -20
