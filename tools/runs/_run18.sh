mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu_weak.json 2> gpurun_out/bench_2gpu_weak.err; echo "weak rc=$?"; tail -3 gpurun_out/bench_2gpu_weak.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --mode strong > gpurun_out/bench_2gpu_strong.json 2> gpurun_out/bench_2gpu_strong.err; echo "strong rc=$?"; tail -3 gpurun_out/bench_2gpu_strong.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/multi_check.py > gpurun_out/multi_check.log 2>&1; echo "multi_check rc=$?"; tail -5 gpurun_out/multi_check.log
python - <<'PY'
import json
for f in ('weak','strong'):
    try:
        d=json.load(open(f'gpurun_out/bench_2gpu_{f}.json')); print(f, round(d['value']), round(d['e2e']['value']), d['ms_per_step'], d['step_wall_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
