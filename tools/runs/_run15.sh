for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench8.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
