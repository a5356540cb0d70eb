# round 2, GPU call J (1 GPU): magic-number tile decode (epilogue instruction diet) -- tests, microbench, sweep; batch-1 latency A/B of the weight hint
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1
tail -4 gpurun_out/r2j_tests.log
timeout 600 python tools/conv_microbench.py --set res4 > gpurun_out/r2j_microbench_res4.txt 2>&1
cat gpurun_out/r2j_microbench_res4.txt
timeout 600 python tools/chunk_sweep.py --steps 20 --out gpurun_out/r2j_sweep.jsonl --config "fastdiv:" --config "fastdiv_again:" > gpurun_out/r2j_sweep.log 2>&1
cut -c1-200 gpurun_out/r2j_sweep.jsonl
for cfg in "filtered DC_L2_HINTS=2" "none DC_L2_HINTS=0" "forced DC_L2_HINTS_SMALL=1" "filtered2 DC_L2_HINTS=2"; do
  set -- $cfg
  env $2 timeout 600 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2j_cfg1_$1.json 2> gpurun_out/r2j_cfg1_$1.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2j_cfg1_$1.json').read()); print('$1', round(d['ms_per_step'],4), 'ms', d['clocks'])"
done
