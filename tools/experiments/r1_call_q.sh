# GPU call Q: split-K exchange: two slots in flight for S = 4; publish without the gpu-scope fence (A/B).
set -x
mkdir -p gpurun_out
for f in 0 1; do
  DC_SK_NOFENCE=$f timeout 120 python tools/conv_microbench.py --set lat 2>&1 | grep latency | sed "s/\$/ nofence=$f/"
done > gpurun_out/q_lat.txt
cat gpurun_out/q_lat.txt
for f in 0 1; do
  DC_SK_NOFENCE=$f timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_net_gpu.py -m gpu -x -q -k "split_k or batch_independence or resnet152 or ragged" 2>&1 | grep -E "passed|failed|Error" | tail -2
  DC_SK_NOFENCE=$f timeout 300 python bench.py --workload cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/q_bench_cfg1_nf$f.json 2> gpurun_out/q_bench_cfg1_nf$f.err
  cut -c1-200 gpurun_out/q_bench_cfg1_nf$f.json; tail -3 gpurun_out/q_bench_cfg1_nf$f.err
done
