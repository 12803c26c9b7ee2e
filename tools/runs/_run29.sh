#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 250 python tools/trace_step.py 8 > gpurun_out/trace8.log 2>&1; grep -A70 "between stitch" gpurun_out/trace8.log | head -150
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_pair.csv python tools/profile_pair.py 5000000 2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
