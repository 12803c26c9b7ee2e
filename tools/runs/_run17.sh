for i in 1 2 3; do
PMN_ALLOC_LOG=1 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench9.json'))
print(round(d['value']), round(d['e2e']['value']), d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
grep -n "timed region" -A6 gpurun_out/bench9.err | head -30
done
