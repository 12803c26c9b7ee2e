timeout 900 python -m pytest tests/test_gpu_post.py -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo rc=$?; tail -3 gpurun_out/bench10.err
python -c "
import json; d=json.load(open('gpurun_out/bench10.json')); print(d['value'], d['e2e']['value'], d['e2e_worker'], d['post_steps_per_pair'], d['device_allocations_in_timed_region'])"
