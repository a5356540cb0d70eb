# GPU call F: scale/shift L1 prefetch + early weight tiles (before griddepcontrol.wait): tests, latency microbench, cfg1 A/B, default bench A/B.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_tests.log 2>&1
tail -5 gpurun_out/f_tests.log
for v in 0 1; do
  DC_EARLY_WEIGHTS=$v timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency | sed "s/\$/ early_weights=$v/"
done > gpurun_out/f_lat.txt
cat gpurun_out/f_lat.txt
for v in 0 1 0 1; do
  DC_EARLY_WEIGHTS=$v timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/f_bench_cfg1_ew$v.json 2> gpurun_out/f_bench_cfg1_ew$v.err
  cut -c1-200 gpurun_out/f_bench_cfg1_ew$v.json; tail -3 gpurun_out/f_bench_cfg1_ew$v.err
done
for v in 0 1; do
  DC_EARLY_WEIGHTS=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/f_bench_n1_ew$v.json 2> gpurun_out/f_bench_n1_ew$v.err
  cut -c1-200 gpurun_out/f_bench_n1_ew$v.json; tail -3 gpurun_out/f_bench_n1_ew$v.err
done
