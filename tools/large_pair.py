"""C4 (one 100 Mbp pair) on the ranks of a torch.distributed.run launch: the index is built on rank 0 and replicated by an
NCCL broadcast of its image, the query positions are sharded for seeding, anchors are all-gathered, clustering and extension
run replicated.  Prints per-phase times (max over ranks, CUDA-synchronised wall clock) as one JSON line on rank 0."""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from paramugsy_b200 import lib, multi, synth
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = synth.config_c4(n=n, inv_len=max(1000, n // 100))
ctx = lib.Context(local)
rs, qs = ctx.sequence(synth.fasta(*g[0])), ctx.sequence(synth.fasta(*g[1]))
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
out = {}
for rep in range(3):
    sync(); t0 = time.perf_counter()
    if rank == 0:
        ix = rs.index()
    else:
        ix = rs.index(empty=True)
    if world > 1:
        dist.broadcast(ix.image_tensor(), src=0)
        torch.cuda.synchronize()
        if rank != 0: ix.adopt()
    sync(); t1 = time.perf_counter()
    res = multi.align_large_pair(ix, qs, rank, world, dist if world > 1 else None, ref_path="c0", qry_path="c1")
    sync(); t2 = time.perf_counter()
    t = torch.tensor([t1 - t0, t2 - t1], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[f"rep{rep}"] = {"index_build_and_broadcast_ms": round(float(t[0]) * 1e3, 2), "align_ms": round(float(t[1]) * 1e3, 2),
                        "delta_sha": hashlib.sha256(res.delta).hexdigest()[:16], "alignments": res.stats["alignments"]}
    res.close(); ix.close()
if rank == 0:
    print(json.dumps({"config": "c4", "bases": n, "gpus": world, **out}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
