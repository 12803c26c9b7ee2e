"""Turns what tools/gpu_visit.sh visits left in gpurun_out/ into the tracked summaries under profiles/.
   python tools/collect_profiles.py <round tag, e.g. r02> <visit tag of the single-GPU pass> [<visit tag> ...]
Copies bench lines (own arm, reference arm, C4, C5, C3, multi-GPU), summarises ncu launch lists (one pair; bench.py itself),
extracts DRAM traffic per launch from the --set full capture and writes a SASS opcode histogram per kernel of the built library."""
import csv, glob, json, os, re, shutil, subprocess, sys
from collections import Counter

R = sys.argv[1]; TAGS = sys.argv[2:]
G, P = "gpurun_out", "profiles"
HERE = os.path.dirname(os.path.abspath(__file__))
os.makedirs(P, exist_ok=True)


def last_json(path):
    try:
        return json.loads(open(path).read().strip().splitlines()[-1])
    except Exception:
        return None


def copy_json(src, dst):
    d = last_json(src)
    if d is not None:
        json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
        return True
    return False


for tag in TAGS:
    for src, dst in ((f"{tag}_bench.json", f"{R}_bench_c2_1gpu.json"), (f"{tag}_bench_ref.json", f"{R}_bench_c2_reference_arm.json"),
                     (f"{tag}_c4_1gpu.json", f"{R}_bench_c4_1gpu.json"), (f"{tag}_c5.json", f"{R}_c5_divergence_sweep.json"), (f"{tag}_c3.json", f"{R}_c3_57x2mbp_job_tree.json")):
        if os.path.exists(os.path.join(G, src)):
            copy_json(os.path.join(G, src), dst)
    for f in glob.glob(os.path.join(G, f"{tag}_c[24]_*gpu*.json")):
        copy_json(f, f"{R}_bench_" + os.path.basename(f)[len(tag) + 1:])
    for f in glob.glob(os.path.join(G, f"{tag}_sweep_*.json")):
        pass
    lp = os.path.join(G, f"{tag}_launches_pair.csv")
    if os.path.exists(lp):
        out = subprocess.run([sys.executable, os.path.join(HERE, "launch_summary.py"), lp, "second_half", os.path.join(P, f"{R}_launches_pair_5mbp_summary.txt")], capture_output=True, text=True)
        head = ("one 5 Mbp pair of C2 (index build + align), second pass of tools/profile_pair.py, ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,"
                "smsp__inst_executed.sum --clock-control none\n")
        t = open(os.path.join(P, f"{R}_launches_pair_5mbp_summary.txt")).read()
        open(os.path.join(P, f"{R}_launches_pair_5mbp_summary.txt"), "w").write(head + t)
    lb = os.path.join(G, f"{tag}_launches_bench.csv")
    if os.path.exists(lb):
        subprocess.run([sys.executable, os.path.join(HERE, "launch_summary.py"), lb, "all", os.path.join(P, f"{R}_launches_bench_summary.txt")], capture_output=True, text=True)
        head = ("bench.py itself (--steps 1 --warmup 1 --no-cpu-baseline: parity pass, warm-up and timed steps of the three arms, the one-pair-at-a-time pass), every launch, "
                "ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none; durations are serialised and cold-cache: shares, not absolutes\n")
        t = open(os.path.join(P, f"{R}_launches_bench_summary.txt")).read()
        open(os.path.join(P, f"{R}_launches_bench_summary.txt"), "w").write(head + t)

# --set full capture of one pair -> per-kernel table and DRAM traffic per launch
rep = os.path.join(G, "pair_full.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.split("\n")))
    hdr, units = rr[0], rr[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
    cols = [w for w in want if w in hdr]
    traffic = {}
    with open(os.path.join(P, f"{R}_ncu_full_pair_5mbp.csv"), "w") as f:
        f.write(",".join(cols) + "\n"); f.write(",".join(units[hdr.index(c)] for c in cols) + "\n")
        for r in rr[2:]:
            if len(r) < len(hdr): continue
            f.write(",".join('"' + r[hdr.index(c)].split("(")[0] + '"' if c == "Kernel Name" else r[hdr.index(c)] for c in cols) + "\n")
            name = r[hdr.index("Kernel Name")].split("(")[0]
            def val(c):
                v = float(r[hdr.index(c)].replace(",", "")); u = units[hdr.index(c)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            t_ns = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1e6, "us": 1e3, "ns": 1, "s": 1e9}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1)
            traffic.setdefault(name, {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "gpu_time_ns": t_ns})
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (first launch of each kernel), one 5 Mbp pair of C2, ncu --set full --clock-control none",
               "kernels": traffic}, open(os.path.join(P, f"{R}_ncu_traffic.json"), "w"), indent=1)

# SASS opcode histogram per kernel of the library as built here
so = os.path.join("paramugsy_b200", "_lib", "libpmnucmer.so")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, hist = None, {}
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); hist[cur] = Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            hist[cur][m.group(1).split(".")[0]] += 1
    with open(os.path.join(P, f"{R}_sass_opcodes.txt"), "w") as f:
        f.write("cuobjdump -sass paramugsy_b200/_lib/libpmnucmer.so (sm_100a): instruction counts per kernel, the 14 most frequent opcodes and every tensor / TMA / DPX opcode\n")
        f.write("no MMA of any kind: nothing on this path is a contraction.  UBLKCP + SYNCS = the TMA bulk copy and its mbarrier in k_seed; VIADDMNMX / VIMNMX3 = DPX min/max forms of the DP\n\n")
        for k in sorted(hist, key=lambda k: -sum(hist[k].values())):
            h = hist[k]; n = sum(h.values())
            if n < 50: continue
            name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
            special = {o: c for o, c in h.items() if re.match(r"(UBLKCP|SYNCS|VIADDMNMX|VIMNMX3|VIMNMX|CREDUX|HMMA|IMMA|UTCMMA|UTMA|REDUX|MATCH|VOTE|SHFL|ATOMS|ATOMG|RED)", o)}
            top = ", ".join(f"{o} {c}" for o, c in h.most_common(14))
            f.write(f"{name}: {n} instructions\n    top: {top}\n    of note: {', '.join(f'{o} {c}' for o, c in sorted(special.items())) or '-'}\n")
print(sorted(os.listdir(P))[-30:])
