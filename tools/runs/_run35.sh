#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for cells in 10000 4096 2500 1600 1024; do
echo "== TPJ_CELLS $cells"
PMN_TPJ_CELLS=$cells timeout 300 python tools/profile_pair.py 5000000 2 2>&1 | tail -1 | grep -o "'ms_extend.*'kernel_launches': [0-9]*"
PMN_TPJ_CELLS=$cells timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bc$cells.json 2>gpurun_out/bc$cells.err; python -c "
import json; d=json.loads(open('gpurun_out/bc$cells.json').read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e_worker']['value'],1), d['step_wall_ms']['resident'])"
done
PMN_TPJ_CELLS=2500 timeout 300 python tools/run_configs.py c5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
for r in d['sweep']: print(r['query'], 'wave1', r['ms_wave1'], 'stitch', r['ms_stitch'], 'total', r['ms_total'])"
