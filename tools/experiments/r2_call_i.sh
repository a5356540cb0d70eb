# round 2, GPU call I (1 GPU): final tree -- all GPU tests, smoke, the bench line + step table, pyramid at N = 1
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1
tail -4 gpurun_out/r2i_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2i_smoke.log 2>&1
tail -3 gpurun_out/r2i_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2i_steps_16x720p.json > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err
cat gpurun_out/r2i_bench_n1.json
timeout 900 python tools/pyramid_bench.py --steps 5 > gpurun_out/r2i_pyramid_n1.json 2> gpurun_out/r2i_pyramid_n1.err
cat gpurun_out/r2i_pyramid_n1.json; tail -3 gpurun_out/r2i_pyramid_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_reference_n1.json 2> gpurun_out/r2i_bench_reference_n1.err
cat gpurun_out/r2i_bench_reference_n1.json
