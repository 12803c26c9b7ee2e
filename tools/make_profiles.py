"""Turns what a tools/gpu_round.sh visit left in gpurun_out/ into the tracked summaries under profiles/."""
import csv, json, os, shutil, subprocess, sys
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
for src, dst in (("bench.json", f"{R}_bench_c2_final.json"), ("bench_ref.json", f"{R}_bench_c2_reference_arm.json"), ("c4.json", f"{R}_c4_100mbp_pair.json"), ("c5.json", f"{R}_c5_divergence_sweep.json"), ("c3.json", f"{R}_c3_57x2mbp_job_tree.json")):
    if os.path.exists(os.path.join(G, src)):
        d = json.load(open(os.path.join(G, src)))
        json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
# launch list: one 5 Mbp pair, second pass (warm)
rows = list(csv.reader(open(os.path.join(G, "launches_pair.csv"))))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]; kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
recs = {}
for r in rows[h + 1:]:
    if len(r) > mv: recs.setdefault(int(r[idc]), {"k": r[kn].split("(")[0]})[r[mn]] = float(r[mv].replace(",", ""))
ids = sorted(recs); half = ids[len(ids) // 2:]
with open(os.path.join(P, f"{R}_launches_pair_5mbp.csv"), "w") as f:
    f.write("id,kernel,gpu_time_ns,sm_cycles_active_avg,warp_inst\n")
    for i in half: r = recs[i]; f.write(f"{i},{r['k']},{r['gpu__time_duration.sum']:.0f},{r['sm__cycles_active.avg']:.0f},{r['smsp__inst_executed.sum']:.0f}\n")
agg = {}
for i in half:
    r = recs[i]; a = agg.setdefault(r["k"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += r["gpu__time_duration.sum"]; a[2] += r["sm__cycles_active.avg"]; a[3] += r["smsp__inst_executed.sum"]
tot = sum(a[1] for a in agg.values()); tota = sum(a[2] for a in agg.values())
with open(os.path.join(P, f"{R}_launches_pair_5mbp_summary.txt"), "w") as f:
    f.write(f"one 5 Mbp pair of C2 (index build + align), second pass, ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --clock-control none\n")
    f.write(f"{len(half)} launches, sum of durations {tot / 1e3:.1f} us (serialised, cold caches), sum of SM-active cycles {tota / 1e3:.0f} kcycles\n")
    f.write(f"{'time us':>10} {'share':>6} {'active kcyc':>12} {'share':>6} {'warp inst M':>12} {'n':>4}  kernel\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{a[1] / 1e3:10.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e3:12.1f} {100 * a[2] / tota:5.1f}% {a[3] / 1e6:12.2f} {a[0]:4d}  {k}\n")
# --set full capture
rep = os.path.join(G, "pair_full.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.split("\n")))
    hdr = rr[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_elapsed.avg"]
    cols = [w for w in want if w in hdr]
    units = rr[1]
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
            traffic.setdefault(name, {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"), "gpu_time_ns": float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")) * {"ms": 1e6, "us": 1e3, "ns": 1, "s": 1e9}.get(units[hdr.index("gpu__time_duration.sum")].lower(), 1)})
    json.dump({"what": "dram__bytes_read.sum + dram__bytes_write.sum per launch, one 5 Mbp pair of C2, ncu --set full --clock-control none", "kernels": traffic}, open(os.path.join(P, f"{R}_ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{R}_launches_pair_5mbp_summary.txt")).read())
print(json.dumps(json.load(open(os.path.join(P, f"{R}_ncu_traffic.json"))), indent=0)[:1500])
