mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_pair.py 5000000 2 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/bench7.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench7.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
timeout 600 python tools/run_configs.py c5 > gpurun_out/c5b.json 2> gpurun_out/c5b.err; echo "c5 rc=$?"; tail -3 gpurun_out/c5b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/c5b.json'))
for r in d['sweep']: print(r['query'], r['ms_total'], r['ms_wave1'], r['ms_stitch'], r['ms_extend'])
PY
