timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
python tools/stage_under_load.py 8 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
