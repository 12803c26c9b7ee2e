"""One 5 Mbp pair of the C2 workload: warm-up pass, then a second pass for ncu (-s skips the first)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
gs = synth.config_c2(n=n, count=2, inv_len=max(1000, n // 100))
ctx = lib.Context(0)
rs, qs = ctx.sequence(synth.fasta(*gs[0])), ctx.sequence(synth.fasta(*gs[1]))
for p in range(passes):
    c0 = ctx.counters()
    ix = rs.index(); res = ix.align(qs); st = res.stats; res.close(); ix.close()
    c1 = ctx.counters()
    print("pass", p, "launches", c1["launches"] - c0["launches"], {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
