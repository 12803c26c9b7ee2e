"""Per-pair stage times (CUDA events) inside a W-worker batch against the same pair alone: which stages stretch under contention."""
import os, sys, statistics as st
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import lib, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = 5_000_000
gs = synth.config_c2(n=n, count=8, inv_len=n // 100)
sched = lib.Scheduler(0, W); ctx = sched.context(0)
seqs = [ctx.sequence(synth.fasta(*g)) for g in gs]; names = [g[0] for g in gs]
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
for _ in range(3):
    for r in sched.align_seqs(seqs, pairs, names=names): r.close()
import time
t = time.perf_counter(); res = sched.align_seqs(seqs, pairs, names=names); wall = (time.perf_counter() - t) * 1e3
keys = ("ms_index", "ms_seed", "ms_seed_kernel", "ms_cluster", "ms_extend", "ms_wave1", "ms_stitch", "ms_total", "wall_ms_align", "wall_ms_text")
print(f"W={W} step wall {wall:.1f} ms")
for k in keys: print(f"  {k:16s} mean {st.mean(r.stats[k] for r in res):8.3f}  max {max(r.stats[k] for r in res):8.3f}  sum {sum(r.stats[k] for r in res):8.1f}")
for r in res: r.close()
