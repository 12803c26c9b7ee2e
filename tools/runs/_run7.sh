mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.avg,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/active_pair.csv python tools/profile_pair.py 5000000 2 > gpurun_out/ncu_active.log 2>&1; echo "rc=$?"
