mkdir -p gpurun_out
timeout 600 python tools/run_configs.py c5 > gpurun_out/c5.json 2> gpurun_out/c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/c5.err
timeout 1200 python tools/run_configs.py c4 > gpurun_out/c4.json 2> gpurun_out/c4.err; echo "c4 rc=$?"; tail -3 gpurun_out/c4.err
cat gpurun_out/c5.json gpurun_out/c4.json
