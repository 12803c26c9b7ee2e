timeout 600 python tools/large_pair.py 1e8 2>/dev/null | tail -1 | tee gpurun_out/c4_1gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/large_pair.py 1e8 2>gpurun_out/c4_2gpu.err | grep '^{' | tail -1 | tee gpurun_out/c4_2gpu.json
tail -3 gpurun_out/c4_2gpu.err
