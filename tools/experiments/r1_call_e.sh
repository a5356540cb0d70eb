# GPU call E: split-K policy (>= 36 K-steps), exit-wait experiment, full tests, BASELINE configs[1] and default bench.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/e_tests.log 2>&1
tail -5 gpurun_out/e_tests.log
for cfg in "1 0" "4 0" "4 1" "1 1"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_EXIT_WAIT_READ=$2 timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency | sed "s/\$/ exit_wait_read=$2/"
done > gpurun_out/e_lat.txt
cat gpurun_out/e_lat.txt
for cfg in "1 0" "4 0" "4 1"; do
  set -- $cfg
  DC_SPLIT_K=$1 DC_EXIT_WAIT_READ=$2 timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_cfg1_sk$1_ewr$2.json 2> gpurun_out/e_bench_cfg1_sk$1_ewr$2.err
  cut -c1-330 gpurun_out/e_bench_cfg1_sk$1_ewr$2.json; tail -3 gpurun_out/e_bench_cfg1_sk$1_ewr$2.err
done
for v in 0 1; do
  DC_EXIT_WAIT_READ=$v timeout 600 python bench.py --no-cpu-baseline > gpurun_out/e_bench_n1_ewr$v.json 2> gpurun_out/e_bench_n1_ewr$v.err
  cut -c1-330 gpurun_out/e_bench_n1_ewr$v.json
done
