# round 2, second 8-GPU call: the exchange leg with SMs reserved for NCCL's kernels / fewer NCCL channels
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
run() { name=$1; shift; env "$@" timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --only-exchange --no-latency-config > gpurun_out/r2n8b_$name.json 2> gpurun_out/r2n8b_$name.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/r2n8b_$name.json').read()); e=d['exchange']; print('$name', 'value', round(d['value'],1), 'exchange', round(e['value'],1), 'ms', round(e['ms_per_step'],2), 'reserved', e['reserved_sms'], e['nccl_env'])"; }
run r8 DC_EXCHANGE_RESERVED_SMS=8
run r0 DC_EXCHANGE_RESERVED_SMS=0
run r16 DC_EXCHANGE_RESERVED_SMS=16
run r8_c4 DC_EXCHANGE_RESERVED_SMS=8 NCCL_MAX_NCHANNELS=4
run r4_c2 DC_EXCHANGE_RESERVED_SMS=4 NCCL_MAX_NCHANNELS=2
run r32 DC_EXCHANGE_RESERVED_SMS=32
