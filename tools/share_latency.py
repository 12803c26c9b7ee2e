"""What a rank of the sharded C2 batch does, timed on ONE GPU: for world = 1, 2, 4, 8 every rank's share of the 28 pairs
(pmn_multi_plan, the plan bench.py --gpus N uses) runs through the scheduler by itself; the step of an N-GPU run is the
slowest share.  A prediction for strong scaling that needs no N-GPU box.  Diagnostics only — never a bench value.
   python tools/share_latency.py [workers] [reps]"""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paramugsy_b200 import lib, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
torch.cuda.init()
gs = synth.config_c2()
fastas = [synth.fasta(*g) for g in gs]
nbytes = [len(f) for f in fastas]
sched = lib.Scheduler(0, W); ctx = sched.context(0)
seqs = [ctx.sequence(f) for f in fastas]; names = [g[0] for g in gs]
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
def run(pl):
    t = time.perf_counter()
    for r in sched.align_seqs(seqs, pl, names=names): r.close()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3
for _ in range(3): run(pairs)
for world in (1, 2, 4, 8):
    dev_of = lib.multi_plan(world, pairs, nbytes)
    worst = 0.0; rows = []
    for r in range(world):
        mine = [p for p, d in zip(pairs, dev_of) if d == r]
        if not mine: continue
        run(mine); run(mine)
        ms = statistics.median(run(mine) for _ in range(reps))
        rows.append((r, len(mine), len({i for i, _ in mine}), ms)); worst = max(worst, ms)
    print(f"world {world}: slowest share {worst:.2f} ms -> {28 / worst * 1e3:.0f} pairs/s predicted; shares (rank, pairs, indexes, ms): " +
          " ".join(f"({r},{n},{k},{ms:.2f})" for r, n, k, ms in rows), flush=True)
one = [statistics.median(run([p]) for _ in range(5)) for p in (pairs[0], pairs[13], pairs[27])]
print("single pairs (index build + pair) ms:", " ".join(f"{x:.2f}" for x in one))
