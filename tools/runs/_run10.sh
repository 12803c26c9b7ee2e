PMN_ALLOC_LOG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
grep -c cudaMalloc gpurun_out/bench6.err; tail -25 gpurun_out/bench6.err
