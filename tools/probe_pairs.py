"""Probe single pairs of the benchmark configurations: stage times and (with PMN_JOBLOG) the
per-engine-call log.  usage: probe_pairs.py c2|c5:<d>|c4:<n>|c1 [passes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import synth, lib
what = sys.argv[1] if len(sys.argv) > 1 else "c2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if what == "c2":
    gs = synth.config_c2(count=2)
elif what == "c1":
    gs = synth.config_c1()
elif what.startswith("c5:"):
    d = float(what[3:]); anc, qs_ = synth.config_c5(ds=(d,)); gs = [anc, qs_[0]]
elif what.startswith("c4:"):
    n = int(float(what[3:])); gs = synth.config_c4(n=n, inv_len=max(1000, n // 100))
ctx = lib.Context(0)
t = time.time()
rs, qs = ctx.sequence(synth.fasta(*gs[0])), ctx.sequence(synth.fasta(*gs[1]))
print(what, "pack wall ms", round((time.time() - t) * 1e3, 2), flush=True)
for p in range(passes):
    if p < passes - 1: os.environ.pop("PMN_JOBLOG", None)
    elif os.environ.get("PMN_JOBLOG_LAST"): os.environ["PMN_JOBLOG"] = os.environ["PMN_JOBLOG_LAST"]
    c0 = ctx.counters(); t = time.time()
    ix = rs.index(); res = ix.align(qs); st = res.stats; n = len(res.delta); res.close(); ix.close()
    c1 = ctx.counters()
    print(what, "pass", p, "wall ms", round((time.time() - t) * 1e3, 2), "launches", c1["launches"] - c0["launches"], "delta bytes", n,
          {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()}, flush=True)
