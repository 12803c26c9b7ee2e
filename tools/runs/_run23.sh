cat > /tmp/maf.py <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
from paramugsy_b200 import lib, synth
gs = synth.config_c2(count=2)
with lib.Context(0) as ctx:
    rs, qs = ctx.sequence(synth.fasta(*gs[0])), ctx.sequence(synth.fasta(*gs[1]))
    ix = rs.index(); res = ix.align(qs); d = res.delta; res.close()
    for k in range(4):
        t = time.perf_counter(); f = ctx.delta_filter(d, 1); t1 = time.perf_counter(); m = ctx.delta2maf(f, rs, qs); t2 = time.perf_counter()
        print(f"filter {1e3*(t1-t):.2f} ms, maf {1e3*(t2-t1):.2f} ms, {len(m)} bytes")
PY
PMN_POST_TIMING=1 python /tmp/maf.py 2>&1 | tail -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:k_maf|k_filter" python /tmp/maf.py 2>&1 | grep -E "k_maf|k_filter|gpu__time" | head -20
