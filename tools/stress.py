"""Stress / sanitizer driver: repeated scheduler batches with many workers.
   python tools/stress.py <genome_bp> <genomes> <workers> <rounds> [post]
Under `compute-sanitizer --tool memcheck` use a small size (e.g. 200000 6 8 2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import lib, synth
n = int(float(sys.argv[1])); g = int(sys.argv[2]); w = int(sys.argv[3]); rounds = int(sys.argv[4]); post = int(sys.argv[5]) if len(sys.argv) > 5 else 0
gs = synth.config_c2(n=n, count=g, inv_len=max(1000, n // 100))
fa = [synth.fasta(*x) for x in gs]
pairs = [(i, j) for i in range(g) for j in range(i + 1, g)]
names = [x[0] + ".fa" for x in gs]
with lib.Scheduler(0, w) as s:
    ctx = s.context(0)
    seqs = [ctx.sequence(f) for f in fa]
    want = None
    t0 = time.time()
    for r in range(rounds):
        for mode in ("seqs", "fasta"):
            res = s.align_seqs(seqs, pairs, names=names) if mode == "seqs" else s.align_fasta(fa, pairs, names=names, post=post)
            got = [x.delta for x in res]
            for x in res: x.close()
            if want is None: want = got
            assert got == want, f"round {r} {mode}: results changed"
        # an index built outside the scheduler on a worker's context, then handed in (the broadcast path of multi.AllVsAll)
        ix = [seqs[k].index() if k < g - 1 else None for k in range(g)]
        res = s.align_seqs(seqs, pairs, names=names, indexes=ix)
        got = [x.delta for x in res]
        for x in res: x.close()
        for x in ix:
            if x is not None: x.close()
        assert got == want, f"round {r} given indexes: results changed"
    print(f"stress ok: {rounds} rounds x 3 batches of {len(pairs)} pairs, {w} workers, {time.time() - t0:.1f} s", flush=True)
    for q in seqs: q.close()
