mkdir -p gpurun_out
nproc; free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/bench_4gpu_weak.json 2> gpurun_out/bench_4gpu_weak.err; echo "weak rc=$?"; tail -2 gpurun_out/bench_4gpu_weak.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 4 --steps 1 --warmup 1 > gpurun_out/bench_4gpu_ref.json 2> gpurun_out/bench_4gpu_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ('weak','ref'):
    try:
        txt=open(f'gpurun_out/bench_4gpu_{f}.json').read(); lines=[l for l in txt.split('\n') if l.strip()]
        print(f, 'lines on stdout:', len(lines))
        d=json.loads(lines[-1]); print(f, round(d['value'],1), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('step_wall_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
