#!/bin/bash
for live in 4 16; do
PMN_SCHED_LIVE_INDEXES=$live timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bl$live.json 2> gpurun_out/bl$live.err; echo "bench rc=$?"
python - $live <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/bl{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('live',sys.argv[1],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'worker',round(d['e2e_worker']['value'],1),d['step_wall_ms']['e2e_worker'])
P
done
