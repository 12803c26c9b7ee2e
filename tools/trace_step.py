"""Kernel timeline of one bench step (28 pairs, W workers) through torch.profiler (CUPTI): how busy the
GPU is, how much the workers' kernels overlap, where host API time goes.  Profiling only — never a bench value."""
import os, sys, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from paramugsy_b200 import lib, synth
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
torch.cuda.init()
gs = synth.config_c2(n=n, count=8, inv_len=max(1000, n // 100))
sched = lib.Scheduler(0, W); ctx = sched.context(0)
seqs = [ctx.sequence(synth.fasta(*g)) for g in gs]; names = [g[0] for g in gs]
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
for _ in range(3):
    for r in sched.align_seqs(seqs, pairs, names=names): r.close()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for r in sched.align_seqs(seqs, pairs, names=names): r.close()
    torch.cuda.synchronize()
out = f"gpurun_out/trace_w{W}.json"
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
k = [e for e in ev if e.get("cat") == "kernel"]
rt = [e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver")]
mc = [e for e in ev if e.get("cat") in ("gpu_memcpy", "gpu_memset")]
t0 = min(e["ts"] for e in k); t1 = max(e["ts"] + e["dur"] for e in k)
print(f"W={W}: {len(k)} kernels, {len(mc)} memcpy/memset, {len(rt)} runtime calls, span {(t1 - t0) / 1e3:.2f} ms")
# union busy time
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in k)
busy = 0; cs, ce = iv[0]
for a, b in iv[1:]:
    if a > ce: busy += ce - cs; cs, ce = a, b
    else: ce = max(ce, b)
busy += ce - cs
print(f"GPU busy (any kernel running) {busy / 1e3:.2f} ms = {100 * busy / (t1 - t0):.1f}% of span; sum of kernel durations {sum(e['dur'] for e in k) / 1e3:.2f} ms")
def union(iv):
    iv = sorted(iv)
    if not iv: return 0
    tot = 0; cs, ce = iv[0]
    for a, b in iv[1:]:
        if a > ce: tot += ce - cs; cs, ce = a, b
        else: ce = max(ce, b)
    return tot + ce - cs
wide = [e for e in k if (e.get("args", {}).get("grid", [1])[0] if isinstance(e.get("args", {}).get("grid"), list) else 1) >= 100]
print(f"wide kernels (grid >= 100 blocks): {len(wide)}, union {union([(e['ts'], e['ts'] + e['dur']) for e in wide]) / 1e3:.2f} ms, sum {sum(e['dur'] for e in wide) / 1e3:.2f} ms")
for nm in ("k_ex_wave1", "k_seed", "k_ex_stitch", "k_cl_chains"):
    sel = [(e['ts'], e['ts'] + e['dur']) for e in k if e["name"].startswith(nm)]
    print(f"  union of {nm}: {union(sel) / 1e3:.2f} ms")
mallocs = [e for e in rt if e["name"] in ("cudaMalloc", "cudaFree", "cudaMallocHost", "cudaFreeHost")]
print("allocation calls in the step:", [(e["name"], round(e["dur"] / 1e3, 2)) for e in mallocs])
agg = collections.defaultdict(lambda: [0, 0.0])
for e in k: a = agg[e["name"].split("(")[0][:50]]; a[0] += 1; a[1] += e["dur"]
for name, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]: print(f"  {name:50s} n={c:5d} total {d / 1e3:8.2f} ms  avg {d / c:8.1f} us")
ra = collections.defaultdict(lambda: [0, 0.0])
for e in rt: a = ra[e["name"]]; a[0] += 1; a[1] += e["dur"]
print("host API:")
for name, (c, d) in sorted(ra.items(), key=lambda kv: -kv[1][1])[:10]: print(f"  {name:40s} n={c:6d} total {d / 1e3:8.2f} ms  avg {d / c:7.1f} us")
os.remove(out)
