# GPU call G: final validation of the round-1 tree + refreshed artefacts for profiles/.
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/g_tests.log 2>&1
tail -4 gpurun_out/g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/g_smoke.log 2>&1; tail -2 gpurun_out/g_smoke.log
( time timeout 900 python bench.py --step-report gpurun_out/g_steps_16x720p.json ) > gpurun_out/g_bench_n1.json 2> gpurun_out/g_bench_n1.err
tail -4 gpurun_out/g_bench_n1.err; cat gpurun_out/g_bench_n1.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --ref-budget-s 45 ) > gpurun_out/g_bench_reference_n1.json 2> gpurun_out/g_bench_reference_n1.err
tail -4 gpurun_out/g_bench_reference_n1.err; cat gpurun_out/g_bench_reference_n1.json
export DC_CUDA_GRAPH=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 324 -c 162 --csv --log-file gpurun_out/g_launches_1x512.csv python tools/profile_forward.py --batch 1 --height 512 --width 512 --warm 2 --iters 1 > gpurun_out/g_ncu_l.log 2>&1
tail -2 gpurun_out/g_ncu_l.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip $((314 + 59)) --launch-count 1 -f -o gpurun_out/g_prof_1x512_res4b7_2b python tools/profile_forward.py --batch 1 --height 512 --width 512 --warm 2 --iters 1 > gpurun_out/g_ncu_sk.log 2>&1
tail -2 gpurun_out/g_ncu_sk.log; ls -la gpurun_out/g_prof_*
