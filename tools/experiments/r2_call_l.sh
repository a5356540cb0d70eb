# round 2, GPU call L (1 GPU): kernels without the removed experiment switches (no spills) -- tests, batch-1 latency, 16x720p step
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1
tail -4 gpurun_out/r2l_tests.log
timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2l_cfg1.json 2> gpurun_out/r2l_cfg1.err
python -c "
import json; d=json.loads(open('gpurun_out/r2l_cfg1.json').read()); print('cfg1', round(d['ms_per_step'],4), 'ms', d['clocks']['sm_mhz'], d['e2e']['value'])"
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2l_sweep.jsonl --config "final:" --config "no_weight_hint:DC_WEIGHTS_EVICT_LAST=0" --config "final_again:" > gpurun_out/r2l_sweep.log 2>&1
cut -c1-130 gpurun_out/r2l_sweep.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 --step-report gpurun_out/r2l_steps_16x720p.json > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err
cat gpurun_out/r2l_bench_n1.json | cut -c1-400
