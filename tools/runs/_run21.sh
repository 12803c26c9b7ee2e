cat > /tmp/q15.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from paramugsy_b200 import lib, synth
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.15
anc, qs_ = synth.config_c5(n=5_000_000, ds=(d,))
with lib.Context(0) as ctx:
    rs = ctx.sequence(synth.fasta(*anc)); ix = rs.index()
    qs = ctx.sequence(synth.fasta(*qs_[0]))
    res = ix.align(qs); print({k: v for k, v in res.stats.items() if k.startswith('ms_') or k in ('dp_cells', 'clusters')}); res.close()
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:k_ex_wave1_big|k_ex_stitch" -c 3 -o gpurun_out/q15_full python /tmp/q15.py 0.12 > gpurun_out/ncu_q15.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_q15.log
