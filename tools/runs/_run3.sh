for w in 2 4 8 12 16 28; do
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --workers $w > gpurun_out/w$w.json 2> gpurun_out/w$w.err
echo "rc=$?"; tail -3 gpurun_out/w$w.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/w$w.json').read()); print('workers',$w,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1), d['step_wall_ms']['resident'], d['device_allocations_in_timed_region'])"
done
