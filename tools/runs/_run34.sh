#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_pair.py 5000000 2 2>&1 | tail -1 | cut -c1-700
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/bx_$name.json 2> gpurun_out/bx_$name.err
  python - $name <<'P'
import json,sys
d=json.loads(open(f'gpurun_out/bx_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1],'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'worker',round(d['e2e_worker']['value'],1))
P
}
run base PMN_SEED_BPS=8
run seed6 PMN_SEED_BPS=6
run seed4 PMN_SEED_BPS=4
run seed6_tpj3 PMN_SEED_BPS=6 PMN_TPJ_BPS=3
run seed6_tpj3_big2 PMN_SEED_BPS=6 PMN_TPJ_BPS=3 PMN_BIG_BPS=2
run seed6_tpj2_big2 PMN_SEED_BPS=6 PMN_TPJ_BPS=2 PMN_BIG_BPS=2
run seed6_big2 PMN_SEED_BPS=6 PMN_BIG_BPS=2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
