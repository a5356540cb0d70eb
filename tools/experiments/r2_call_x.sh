# round 2, GPU call X (1 GPU): the default bench line of the committed final tree (all records), as the driver will run it
set -x
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err
python -c "
import json; d=json.loads(open('gpurun_out/r2x_bench_default.json').read()); print(sorted(d.keys())); print(d['ms_per_step'], d['value'], d['e2e'], d['gpu_launches'], d['clocks'])"
tail -3 gpurun_out/r2x_bench_default.err
