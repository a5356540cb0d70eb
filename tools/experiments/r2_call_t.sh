# round 2, GPU call T (2 GPUs): the world-2 NCCL exchange test and the N = 2 bench line (exchange record) on the final tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_exchange_gpu.py tests/test_pyramid_gpu.py -x -q > gpurun_out/r2t_tests.log 2>&1
tail -4 gpurun_out/r2t_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
  > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/r2t_bench_n2.json').read()); print(d['value'], d['e2e']['value'], d['exchange']['value'] if d.get('exchange') else None, d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/pyramid_bench.py > gpurun_out/r2t_pyramid_n2.json 2> gpurun_out/r2t_pyramid_n2.err
cut -c1-300 gpurun_out/r2t_pyramid_n2.json
