mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_pair.py 5000000 2 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; tail -3 gpurun_out/bench4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench4.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'], d['roofline_extend'])
PY
