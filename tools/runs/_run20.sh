for w in 6 10 12 16; do
timeout 600 python bench.py --steps 10 --warmup 4 --no-cpu-baseline --workers $w > gpurun_out/bw$w.json 2> gpurun_out/bw$w.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bw$w.json'))
print('workers',$w, round(d['value']), round(d['e2e']['value']), d['step_wall_ms']['resident'], d['device_allocations_in_timed_region'])
PY
done
