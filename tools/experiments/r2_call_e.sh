# round 2, GPU call E (2 GPUs): NCCL exchange path -- world-2 test, bench line with the `exchange` record, pyramid on 2 ranks
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2e_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_exchange_gpu.py -x -q > gpurun_out/r2e_test_exchange.log 2>&1
tail -15 gpurun_out/r2e_test_exchange.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
cat gpurun_out/r2e_bench_n2.json; tail -5 gpurun_out/r2e_bench_n2.err
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --workload batch128_512 > gpurun_out/r2e_bench_cfg3_n2.json 2> gpurun_out/r2e_bench_cfg3_n2.err
cat gpurun_out/r2e_bench_cfg3_n2.json
timeout 900 $TR tools/pyramid_bench.py --steps 5 > gpurun_out/r2e_pyramid_n2.json 2> gpurun_out/r2e_pyramid_n2.err
cat gpurun_out/r2e_pyramid_n2.json; tail -5 gpurun_out/r2e_pyramid_n2.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2e_bench_reference_n2.json 2> gpurun_out/r2e_bench_reference_n2.err
cat gpurun_out/r2e_bench_reference_n2.json
