# round 2, GPU call N (1 GPU): new tests (skipped outputs, survey-recipe weights), concurrent sub-batch streams experiment,
# what bounds the 1x1 convs now (operand-skip microbenchmark on the elect build), bench line with the without_next_pred record,
# ncu launch list + full captures of one res4 block (2a / 2b / 2c) on the elect build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1
tail -4 gpurun_out/r2n_tests.log
grep -h "parity\] survey" gpurun_out/r2n_tests.log
timeout 600 python tools/stream_split_bench.py --steps 20 --splits 1,2,4,1 --out gpurun_out/r2n_stream_split.jsonl > gpurun_out/r2n_stream_split.log 2>&1
cat gpurun_out/r2n_stream_split.jsonl
timeout 300 python tools/conv_microbench.py --set bound > gpurun_out/r2n_micro_bound.txt 2>&1
cat gpurun_out/r2n_micro_bound.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_n1.json 2> gpurun_out/r2n_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2n_bench_n1.json').read()); print(d['ms_per_step'], d['without_next_pred'], d['clocks'])"
export DC_CUDA_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2n_launches_16x720p.csv python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2n_ncu_l.log 2>&1
for spec in res4b7_2a:58 res4b7_2b:59 res4b7_2c:60; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip $idx --launch-count 1 -f -o gpurun_out/r2n_prof_$name \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2n_ncu_$name.log 2>&1
done
ls -la gpurun_out | tail -12
