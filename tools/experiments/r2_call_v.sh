# round 2, GPU call V (8 GPUs): the N = 8 bench line (device-resident, e2e, NCCL exchange record) on the final tree
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/r2v_bench_n8.json').read()); print(d['value'], d['e2e']['value'], d['exchange']['value'] if d.get('exchange') else None, d['clocks'])"
