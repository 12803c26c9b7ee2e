"""Where do the rare long e2e steps of bench.py come from?  Runs the e2e arm (FASTA text in pinned host memory -> .delta text) many
times in one process under three conditions: nvidia-smi sampling as bench.py does it (clocks, reasons and power.draw every 200 ms),
the same query without power.draw, and no sampler at all; prints how many steps took more than 1.3x the median.
usage: outlier_probe.py [steps per condition]"""
import os, statistics, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from paramugsy_b200 import lib, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
torch.cuda.init()
gs = synth.config_c2()
names = [g[0] for g in gs]
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
pinned = [torch.frombuffer(bytearray(synth.fasta(*g)), dtype=torch.uint8).pin_memory() for g in gs]
fasta_bytes = [(t.data_ptr(), t.numel()) for t in pinned]
sched = lib.Scheduler(0, 32)
def step(post=0):
    t = time.perf_counter()
    for r in sched.align_fasta(fasta_bytes, pairs, names=names, post=post): r.close()
    return (time.perf_counter() - t) * 1e3
for _ in range(5): step(); step(1)
Q_FULL = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
Q_NOPOWER = Q_FULL.replace("power.draw,", "")
Q_CLOCKS = "index,clocks.sm,clocks.max.sm"
def with_sampler(q, post):
    proc = None
    if q:
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", "0"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        rows = []
        threading.Thread(target=lambda: [rows.append(l) for l in proc.stdout], daemon=True).start()
        t = time.time()
        while not rows and time.time() - t < 15: time.sleep(0.05)
    w = [step(post) for _ in range(steps)]
    if proc: proc.terminate(); proc.wait()
    med = statistics.median(w)
    out = sorted(x for x in w if x > 1.3 * med)
    return med, len(out), [round(x, 1) for x in out[-6:]]
for post in (0, 1):
    for name, q in (("bench.py's query (with power.draw)", Q_FULL), ("without power.draw", Q_NOPOWER), ("clocks only", Q_CLOCKS), ("no sampler", None)):
        med, n, worst = with_sampler(q, post)
        print(f"post={post} {name:36s}: median {med:6.2f} ms, {n:3d} of {steps} steps above 1.3x median, worst {worst}", flush=True)
