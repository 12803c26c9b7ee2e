#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python tools/run_configs.py c3 > gpurun_out/c3.json 2> gpurun_out/c3.err; echo "c3 rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; head -c 300 gpurun_out/bench.json; echo; cat gpurun_out/c3.json | head -c 400
