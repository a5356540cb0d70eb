# GPU call A (round 1, session 2): validate the tree, then A/B the small-grid BN=64 heuristic on BASELINE configs[1].
set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/a_tests.log 2>&1
tail -25 gpurun_out/a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -3 gpurun_out/a_smoke.log
for v in 0 1; do
  DC_SMALL_GRID_BN64=$v timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline --step-report gpurun_out/a_steps_cfg1_bn64_$v.json > gpurun_out/a_bench_cfg1_bn64_$v.json 2> gpurun_out/a_bench_cfg1_bn64_$v.err
  cat gpurun_out/a_bench_cfg1_bn64_$v.json
done
timeout 600 python bench.py --no-cpu-baseline --step-report gpurun_out/a_steps_16x720p.json > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
cat gpurun_out/a_bench_n1.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
