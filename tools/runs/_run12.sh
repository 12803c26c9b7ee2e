for ks in 0 1; do
PMN_INDEX_KSHIFT=$ks timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_k$ks.json 2> gpurun_out/bench_k$ks.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_k$ks.json'))
print('kshift',$ks, d['value'], d['e2e']['value'], d['step_wall_ms']['resident'], d['device_allocations_in_timed_region'], d['stage_ms_per_step'])
PY
done
