#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in 8 12 14; do
timeout 300 python bench.py --workers $w --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bw$w.json 2> gpurun_out/bw$w.err; echo "bench w=$w rc=$?"
python - $w <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/bw{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('W',sys.argv[1],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'worker',round(d['e2e_worker']['value'],1),'stage',d['stage_ms_per_step'])
P
done
timeout 250 python tools/trace_step.py 8 > gpurun_out/trace8.log 2>&1; grep -A20 "per stream" gpurun_out/trace8.log
timeout 300 python tools/profile_pair.py 5000000 2 > gpurun_out/pair.log 2>&1; tail -1 gpurun_out/pair.log
