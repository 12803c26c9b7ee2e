#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_pair.py 5000000 2 > gpurun_out/pair.log 2>&1; tail -1 gpurun_out/pair.log
timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bw8.json 2> gpurun_out/bw8.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/bw8.json').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'worker',round(d['e2e_worker']['value'],1),'stage',d['stage_ms_per_step'])
P
timeout 600 python tools/run_configs.py c5 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c5.json'))
for r in d['sweep']: print(r['query'], r['clusters'], 'wave1', r['ms_wave1'], 'stitch', r['ms_stitch'], 'extend', r['ms_extend'], 'total', r['ms_total'])"
