"""One pair of the C5 sweep (default 10 % divergence): warm-up pass, then a second pass for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import synth, lib
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.10
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
anc, qs_ = synth.config_c5(n=n, ds=(d,))
ctx = lib.Context(0)
rs = ctx.sequence(synth.fasta(*anc)); ix = rs.index()
qs = ctx.sequence(synth.fasta(*qs_[0]))
for p in range(2):
    res = ix.align(qs); st = res.stats; res.close()
    print("pass", p, {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
