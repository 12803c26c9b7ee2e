mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_ex_wave1|k_ex_stitch" -c 3 -o gpurun_out/ext_full python tools/profile_pair.py 5000000 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
