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
E2E = len(sys.argv) > 3 and sys.argv[3] == "e2e"       # FASTA text in pinned host memory -> .delta text (what bench.py's e2e arm times)
pinned = [torch.frombuffer(bytearray(synth.fasta(*g)), dtype=torch.uint8).pin_memory() for g in gs] if E2E else []
fasta_bytes = [(t.data_ptr(), t.numel()) for t in pinned]
def step():
    res = sched.align_fasta(fasta_bytes, pairs, names=names) if E2E else sched.align_seqs(seqs, pairs, names=names)
    for r in res: r.close()
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
out = f"gpurun_out/trace_w{W}.json"
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
k = [e for e in ev if e.get("cat") == "kernel"]
rt = [e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver")]
mc = [e for e in ev if e.get("cat") in ("gpu_memcpy", "gpu_memset")]
t0 = min(e["ts"] for e in k + mc); t1 = max(e["ts"] + e["dur"] for e in k + mc)
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
if E2E:
    big = sorted((e for e in mc if e.get("args", {}).get("bytes", 0) >= 1 << 20), key=lambda e: e["ts"])
    print("copies of 1 MB and more (start ms, ms, MB, GB/s, stream):", [(round((e["ts"] - t0) / 1e3, 2), round(e["dur"] / 1e3, 2), round(e["args"]["bytes"] / 2**20, 1), round(e["args"]["bytes"] / e["dur"] / 1e3, 1), e["args"].get("stream")) for e in big])
    fa = sorted((e for e in k if e["name"].startswith("k_pack") or e["name"].startswith("k_revcomp")), key=lambda e: e["ts"])
    print("k_pack / k_revcomp ends (ms):", [round((e["ts"] + e["dur"] - t0) / 1e3, 2) for e in fa])
    print("first copy or kernel at", round((min(e["ts"] for e in k + mc) - t0) / 1e3, 2), "ms relative to the first kernel")
mallocs = [e for e in rt if e["name"] in ("cudaMalloc", "cudaFree", "cudaMallocHost", "cudaFreeHost")]
print("allocation calls in the step:", [(e["name"], round(e["dur"] / 1e3, 2)) for e in mallocs])
agg = collections.defaultdict(lambda: [0, 0.0])
for e in k: a = agg[e["name"].split("(")[0][:50]]; a[0] += 1; a[1] += e["dur"]
for name, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]: print(f"  {name:50s} n={c:5d} total {d / 1e3:8.2f} ms  avg {d / c:8.1f} us")
ra = collections.defaultdict(lambda: [0, 0.0])
for e in rt: a = ra[e["name"]]; a[0] += 1; a[1] += e["dur"]
print("host API:")
for name, (c, d) in sorted(ra.items(), key=lambda kv: -kv[1][1])[:10]: print(f"  {name:40s} n={c:6d} total {d / 1e3:8.2f} ms  avg {d / c:7.1f} us")
# per-stream (= per-worker) view: when each stream ran which pair phase, and how much of the span its stream was busy
by = collections.defaultdict(list)
for e in k: by[e.get("args", {}).get("stream", 0)].append(e)
print("per stream: first kernel, last kernel, busy (union) ms, kernels; then the start times (ms) of k_sa_keys [I], k_seed [S], k_ex_stitch [X]")
for sid, evs in sorted(by.items()):
    evs.sort(key=lambda e: e["ts"])
    marks = []
    for e in evs:
        nm = e["name"]
        if nm.startswith("k_sa_keys"): marks.append(f"I{(e['ts'] - t0) / 1e3:.1f}")
        elif nm.startswith("k_seed("): marks.append(f"S{(e['ts'] - t0) / 1e3:.1f}")
        elif nm.startswith("k_ex_stitch"): marks.append(f"X{(e['ts'] + e['dur'] - t0) / 1e3:.1f}")
    print(f"  stream {sid}: {(evs[0]['ts'] - t0) / 1e3:6.2f} .. {(evs[-1]['ts'] + evs[-1]['dur'] - t0) / 1e3:6.2f}  busy {union([(e['ts'], e['ts'] + e['dur']) for e in evs]) / 1e3:6.2f}  n={len(evs)}  " + " ".join(marks))
# what one worker does between the end of its first stitch and its next seed launch
sid0 = sorted(by)[0]
evs = sorted(by[sid0], key=lambda e: e["ts"])
xs = [e for e in evs if e["name"].startswith("k_ex_stitch")]
if xs:
    xa = xs[0]["ts"] + xs[0]["dur"]
    nxt = [e for e in evs if e["name"].startswith("k_seed(") and e["ts"] > xa]
    xb = nxt[0]["ts"] if nxt else xa + 3000
    print(f"stream {sid0} between stitch end {(xa - t0) / 1e3:.2f} and next seed {(xb - t0) / 1e3:.2f} ms (GPU side):")
    for e in sorted(k + mc, key=lambda e: e["ts"]):
        if e.get("args", {}).get("stream", 0) == sid0 and xa - 50 <= e["ts"] <= xb:
            print(f"   +{(e['ts'] - xa):8.1f} us  dur {e['dur']:7.1f}  {e['name'][:60]}  {e.get('args', {}).get('bytes', '')}")
    # the host thread that launched the stitch
    corr = xs[0].get("args", {}).get("correlation")
    tid = None
    for e in rt:
        if e.get("args", {}).get("correlation") == corr: tid = e.get("tid")
    print(f"host thread {tid} in the same window:")
    for e in sorted(rt, key=lambda e: e["ts"]):
        if e.get("tid") == tid and xa - 50 <= e["ts"] <= xb and (e["dur"] > 15 or e["name"] != "cudaLaunchKernel"):
            print(f"   +{(e['ts'] - xa):8.1f} us  dur {e['dur']:7.1f}  {e['name']}")
# inside one pair (second seed of the first stream .. its stitch end): where the stream sat idle
seeds = [e for e in evs if e["name"].startswith("k_seed(")]
if len(seeds) >= 2 and len(xs) >= 2:
    a = seeds[1]["ts"]; xe = [e for e in xs if e["ts"] > a]
    b = xe[0]["ts"] + xe[0]["dur"] if xe else a
    act = sorted([e for e in k + mc if e.get("args", {}).get("stream", 0) in (sid0, sid0 + 1) and a <= e["ts"] <= b], key=lambda e: e["ts"])
    busy_in = union([(e["ts"], e["ts"] + e["dur"]) for e in act])
    print(f"pair window {(b - a) / 1e3:.2f} ms, stream busy {busy_in / 1e3:.2f} ms, {len(act)} activities; idle gaps > 20 us:")
    end = act[0]["ts"] + act[0]["dur"]; prev = act[0]
    for e in act[1:]:
        if e["ts"] - end > 20: print(f"   gap {e['ts'] - end:7.1f} us at +{(end - a) / 1e3:6.3f} ms after {prev['name'][:36]:36s} before {e['name'][:36]}")
        if e["ts"] + e["dur"] > end: end = e["ts"] + e["dur"]; prev = e
print("host cores:", os.cpu_count())
# concurrency histogram in 1 ms bins: kernels running, wide kernels running (time-weighted averages)
nb = int((t1 - t0) / 1e3) + 1
run = [0.0] * nb; wrun = [0.0] * nb
wide_ids = set(id(e) for e in wide)
for e in k:
    a, b = e["ts"] - t0, e["ts"] + e["dur"] - t0
    for bi in range(int(a / 1e3), min(nb - 1, int(b / 1e3)) + 1):
        ov = max(0.0, min(b, (bi + 1) * 1e3) - max(a, bi * 1e3)) / 1e3
        run[bi] += ov
        if id(e) in wide_ids: wrun[bi] += ov
print("avg kernels running per 1 ms bin:", " ".join(f"{x:.1f}" for x in run))
print("avg wide kernels running per 1 ms bin:", " ".join(f"{x:.1f}" for x in wrun))
os.remove(out)
