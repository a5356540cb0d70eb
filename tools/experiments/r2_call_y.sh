# round 2, GPU call Y (1 GPU): bench line after the arena / weights reporting fix
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2y_bench.json').read()); print(d['ms_per_step'], d['arena_mib'], d['weights_mib'], d['e2e']['device_mib_per_net'], d['e2e']['nets_per_rank'], d['without_next_pred']['ms_per_step'])"
tail -3 gpurun_out/r2y_bench.err
