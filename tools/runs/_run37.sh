#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -i "^CPU(s)\|NUMA node(s)\|Model name" | head -4
run() { name=$1; shift; port=$1; shift
 env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/b8_$name.json 2> gpurun_out/b8_$name.err
 python - $name <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/b8_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], round(d['value'],1), round(d['e2e']['value'],1), 'cores', d.get('host_cores'), d['ms_per_step_by_rank'])
P
}
EXTRA="" run w8 29541 A=1
EXTRA="--workers 6" run w6 29542 A=1
EXTRA="" run w8_yield 29543 PMN_DEVICE_SCHED=yield
