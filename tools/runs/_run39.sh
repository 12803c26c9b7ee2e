#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/b8_default.json 2> gpurun_out/b8_default.err; echo "rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/b8_default.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['e2e']['value'],1), 'W', d['config']['workers_per_gpu'], 'cores', d.get('host_cores'), d['ms_per_step_by_rank'])
P
