# round 2, 8-GPU call: the driver's scaling invocation (device-resident value, per-rank e2e, NCCL exchange record), configs[3], configs[4]
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n8_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2n8_bench_n8.json 2> gpurun_out/r2n8_bench_n8.err
cat gpurun_out/r2n8_bench_n8.json; tail -3 gpurun_out/r2n8_bench_n8.err
timeout 900 $TR bench.py --gpus 8 --steps 20 --warmup 5 --workload batch128_512 > gpurun_out/r2n8_bench_batch128_512_n8.json 2> gpurun_out/r2n8_bench_batch128_512_n8.err
cat gpurun_out/r2n8_bench_batch128_512_n8.json
timeout 900 $TR tools/pyramid_bench.py --steps 8 > gpurun_out/r2n8_pyramid_n8.json 2> gpurun_out/r2n8_pyramid_n8.err
cat gpurun_out/r2n8_pyramid_n8.json; tail -3 gpurun_out/r2n8_pyramid_n8.err
