timeout 900 python tools/run_configs.py c3 > gpurun_out/c3.json 2> gpurun_out/c3.err; echo "c3 rc=$?"; tail -3 gpurun_out/c3.err; cat gpurun_out/c3.json
