"""Multi-GPU checks, launched with torch.distributed.run (one rank per GPU):
  all-vs-all strong mode == single-GPU results; a large pair sharded over the ranks == undivided run."""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from paramugsy_b200 import lib, multi, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000
big = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
sched = lib.Scheduler(local, 4); ctx = sched.context(0)
gs = synth.config_c2(n=n, count=6, inv_len=max(1000, n // 100))
names = [g[0] for g in gs]; seqs = [ctx.sequence(synth.fasta(*g)) for g in gs]
pairs = [(i, j) for i in range(6) for j in range(i + 1, 6)]
single = {k: r.delta for k, r in enumerate(sched.align_seqs(seqs, pairs, names=names))}
ava = multi.AllVsAll(sched, seqs, names, pairs, rank, world, dist)
for it in range(2):
    out = ava.step()
    assert sorted(out) == sorted(ava.mine)
    for k, r in out.items():
        assert r.delta == single[k], f"rank {rank}: pair {k} differs"
counts = torch.tensor([len(ava.mine)], device="cuda"); dist.all_reduce(counts)
assert int(counts.item()) == len(pairs)
print(f"rank {rank}: all-vs-all ok, {len(ava.mine)} pairs, plan {ava.plan}", flush=True)
# one large pair
g4 = synth.config_c4(n=big, inv_len=max(1000, big // 100))
rs, qs = ctx.sequence(synth.fasta(*g4[0])), ctx.sequence(synth.fasta(*g4[1]))
ix = rs.index()
whole = ix.align(qs, ref_path="c0", qry_path="c1").delta
torch.cuda.synchronize(); dist.barrier(); t = time.time()
res = multi.align_large_pair(ix, qs, rank, world, dist, ref_path="c0", qry_path="c1")
torch.cuda.synchronize(); dt = time.time() - t
assert res.delta == whole, f"rank {rank}: sharded large pair differs"
h = hashlib.sha256(res.delta).hexdigest()[:16]
print(f"rank {rank}: large pair ok ({len(whole)} delta bytes, sha {h}, {dt * 1e3:.1f} ms sharded)", flush=True)
dist.barrier(); dist.destroy_process_group()
