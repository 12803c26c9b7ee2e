"""Timeline of ONE C2 pair alone (index resident) through torch.profiler (CUPTI): every kernel, copy and memset of the
alignment on the GPU with its start and duration, the gaps between them, and the host calls that block.
Profiling only — never a bench value.  usage: trace_pair.py [bp] [pair|index]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from paramugsy_b200 import lib, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 5_000_000
what = sys.argv[2] if len(sys.argv) > 2 else "pair"
torch.cuda.init()
gs = synth.config_c2(n=n, count=2, inv_len=max(1000, n // 100))
ctx = lib.Context(0)
rs, qs = ctx.sequence(synth.fasta(*gs[0])), ctx.sequence(synth.fasta(*gs[1]))
ix = rs.index()
for _ in range(3):
    ix.align(qs).close()
torch.cuda.synchronize()
for _ in range(2):
    rs.index().close()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    if what == "index":
        rs.index().close(); st = {k: 0.0 for k in ("ms_seed", "ms_cluster", "ms_extend", "ms_wave1", "ms_stitch", "ms_total")}
    else:
        res = ix.align(qs); st = res.stats; res.close()
    torch.cuda.synchronize()
out = "gpurun_out/trace_pair.json"
prof.export_chrome_trace(out)
ev = json.load(open(out))["traceEvents"]
os.remove(out)
g = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")), key=lambda e: e["ts"])
t0 = g[0]["ts"]
print(f"{len(g)} GPU operations, span {(g[-1]['ts'] + g[-1]['dur'] - t0) / 1e3:.3f} ms, sum of durations {sum(e['dur'] for e in g) / 1e3:.3f} ms; stats: "
      + ", ".join(f"{k} {st[k]:.3f}" for k in ("ms_seed", "ms_cluster", "ms_extend", "ms_wave1", "ms_stitch", "ms_total")))
print("   start us   dur us   gap us  stream  operation")
end = t0
for e in g:
    gap = e["ts"] - end
    print(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} {gap:8.1f}  {e.get('args', {}).get('stream', '?'):>6}  {e['name'][:70]}")
    end = max(end, e["ts"] + e["dur"])
rt = sorted((e for e in ev if e.get("cat") in ("cuda_runtime", "cuda_driver") and e["dur"] >= 15 and t0 - 200 <= e["ts"] <= end), key=lambda e: e["ts"])
print("host calls of 15 us and more:")
for e in rt:
    print(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f}  {e['name']}")
