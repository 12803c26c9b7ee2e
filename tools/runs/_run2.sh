mkdir -p gpurun_out
PMN_ALLOC_LOG=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"
timeout 600 python tools/trace_step.py 8 > gpurun_out/trace8.log 2>&1; echo "trace rc=$?"
timeout 600 python tools/trace_step.py 16 > gpurun_out/trace16.log 2>&1; echo "trace rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print(d['value'], d['e2e']['value'], d['step_wall_ms'], d['device_allocations_in_timed_region'])
PY
tail -30 gpurun_out/bench2.err
cat gpurun_out/trace8.log
