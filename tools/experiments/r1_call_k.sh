# GPU call K: activation-map L2 promotion A/B on the throughput workload, then the final validation of the round-1 tree.
set -x
mkdir -p gpurun_out
for v in 128 256 128 256; do
  DC_ACT_L2_PROMO=$v timeout 600 python bench.py --no-cpu-baseline --no-latency-config > gpurun_out/k_bench_n1_promo$v.json 2> gpurun_out/k_bench_n1_promo$v.err
  python -c "import json;d=json.load(open('gpurun_out/k_bench_n1_promo$v.json'));print('promo $v', d['value'], d['ms_per_step'], d['clocks'])"
done
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/k_tests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/k_tests.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/k_smoke.log 2>&1; tail -2 gpurun_out/k_smoke.log
( time timeout 900 python bench.py --step-report gpurun_out/k_steps_16x720p.json ) > gpurun_out/k_bench_n1.json 2> gpurun_out/k_bench_n1.err
tail -4 gpurun_out/k_bench_n1.err; cat gpurun_out/k_bench_n1.json
