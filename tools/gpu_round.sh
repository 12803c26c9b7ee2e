#!/bin/bash
# One GPU visit: parity tests, smoke, bench (own arm + reference arm), the other configs, ncu launch list and --set full captures.
mkdir -p gpurun_out
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/run_configs.py c5 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"
timeout 900 python tools/run_configs.py c4 > gpurun_out/c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"
timeout 900 python tools/run_configs.py c3 > gpurun_out/c3.json 2> gpurun_out/c3.err; echo "c3 rc=$?"
timeout 600 python tools/profile_pair.py 5000000 2 > gpurun_out/pair.log 2>&1; echo "pair rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/launches_pair.csv python tools/profile_pair.py 5000000 2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_ex_wave|k_ex_stitch|k_seed$|k_cl_chains|k_lcp|k_sa_keys|pmn_rs_scatter|k_bucket_fill" -c 16 -o gpurun_out/pair_full python tools/profile_pair.py 5000000 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; head -c 600 gpurun_out/bench.json
