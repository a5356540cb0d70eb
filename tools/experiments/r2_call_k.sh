# round 2, GPU call K (1 GPU): does the .L2::cache_hint instruction form itself cost latency at batch 1?  A/B by rebuilding on the box
set -x
mkdir -p gpurun_out
lat() { timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2k_cfg1_$1.json 2> gpurun_out/r2k_cfg1_$1.err; python -c "
import json; d=json.loads(open('gpurun_out/r2k_cfg1_$1.json').read()); print('$1', round(d['ms_per_step'],4), 'ms', d['clocks']['sm_mhz'])"; }
lat hinted_a
DC_EXTRA_NVCC_FLAGS=-DDC_PTX_NO_CACHE_HINT python -c "
import importlib; b=importlib.import_module('deepcut-cnn_b200.build'); b.build_kernels(force=True); b.build_host(force=True)"
lat plain_a
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2k_sweep.jsonl --config "plain_build:" > gpurun_out/r2k_sweep_a.log 2>&1
python -c "
import importlib; b=importlib.import_module('deepcut-cnn_b200.build'); b.build_kernels(force=True); b.build_host(force=True)"
lat hinted_b
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2k_sweep.jsonl --config "hinted_build:" --config "hinted_build_nohints:DC_L2_HINTS=0" > gpurun_out/r2k_sweep_b.log 2>&1
cut -c1-160 gpurun_out/r2k_sweep.jsonl
