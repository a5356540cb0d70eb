# round 2, GPU call M (1 GPU): TMA / MMA issue by an elect.sync lane of a converged warp (no ptxas waterfall) vs the lane-0 form,
# same box: tests on the new build, per-layer microbenchmark + 16x720p step for both builds, then the bench line of the new build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1
tail -4 gpurun_out/r2m_tests.log
timeout 300 python tools/conv_microbench.py --set pairs > gpurun_out/r2m_micro_elect.txt 2>&1
cat gpurun_out/r2m_micro_elect.txt
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2m_sweep.jsonl --config "elect:" > gpurun_out/r2m_sweep_a.log 2>&1
DC_EXTRA_NVCC_FLAGS=-DDC_ISSUE_LANE0 python -c "
import importlib; b=importlib.import_module('deepcut-cnn_b200.build'); b.build_kernels(force=True); b.build_host(force=True)"
timeout 300 python tools/conv_microbench.py --set pairs > gpurun_out/r2m_micro_lane0.txt 2>&1
cat gpurun_out/r2m_micro_lane0.txt
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2m_sweep.jsonl --config "lane0:" > gpurun_out/r2m_sweep_b.log 2>&1
python -c "
import importlib; b=importlib.import_module('deepcut-cnn_b200.build'); b.build_kernels(force=True); b.build_host(force=True)"
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2m_sweep.jsonl --config "elect_again:" > gpurun_out/r2m_sweep_c.log 2>&1
cut -c1-200 gpurun_out/r2m_sweep.jsonl
timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2m_cfg1.json 2> gpurun_out/r2m_cfg1.err
cut -c1-300 gpurun_out/r2m_cfg1.json
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2m_steps_16x720p.json > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
cut -c1-400 gpurun_out/r2m_bench_n1.json
