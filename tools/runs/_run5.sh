mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pair.csv python tools/profile_pair.py 5000000 2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
PMN_JOBLOG=gpurun_out/joblog.txt timeout 300 python tools/profile_pair.py 5000000 1 > /dev/null 2>&1
python tools/joblog_summary.py gpurun_out/joblog.txt 2>&1 | tail -40
