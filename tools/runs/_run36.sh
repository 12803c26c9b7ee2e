#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_weak.json 2> gpurun_out/bench_8gpu_weak.err; echo "weak rc=$?"; tail -2 gpurun_out/bench_8gpu_weak.err; tail -c 1500 gpurun_out/bench_8gpu_weak.json | head -c 600
