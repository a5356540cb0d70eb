# round 2, GPU call S (1 GPU): evidence for the final tree -- bench line (with the CPU baseline leg) + step table, batch-1 latency with and
# without TALL mode, ncu launch list of one forward, ncu --set full of the stem conv and res2b's 3x3 (TALL) and of one res4 block
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2s_steps_16x720p.json > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_n1.json').read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['without_next_pred'], d['clocks'], d['roofline']['frac'], d['roofline']['issued_frac'], d['cpu_baseline'], d['latency_config'])"
for v in 1 0 1 0; do
DC_CONV_TALL=$v timeout 600 python bench.py --workload cfg1 --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/r2s_cfg1_tall$v.json 2> gpurun_out/r2s_cfg1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s_cfg1_tall$v.json').read()); print('cfg1 tall=$v', d['ms_per_step'], d['e2e']['value'])"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_bench_reference_n1.json 2> gpurun_out/r2s_bench_reference_n1.err
cut -c1-400 gpurun_out/r2s_bench_reference_n1.json
export DC_CUDA_GRAPH=0
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/r2s_launches_16x720p.csv python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2s_ncu_l.log 2>&1
for spec in conv1:0 res2b_2b:6 res4b7_2a:58 res4b7_2b:59 res4b7_2c:60; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm --launch-skip $idx --launch-count 1 -f -o gpurun_out/r2s_prof_$name \
     python tools/profile_forward.py --warm 2 --iters 1 --profiler-range > gpurun_out/r2s_ncu_$name.log 2>&1
done
ls -la gpurun_out | grep r2s
