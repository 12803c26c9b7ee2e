"""Summarises an ncu launch list (--metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum --csv):
   python tools/launch_summary.py <launches.csv> [second_half|all] [out.txt]
second_half: only the second half of the launches (the warm pass of tools/profile_pair.py)."""
import csv, sys
src = sys.argv[1]; which = sys.argv[2] if len(sys.argv) > 2 else "second_half"
rows = list(csv.reader(open(src)))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[h]; kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
un = hdr.index("Metric Unit")
recs = {}
for r in rows[h + 1:]:
    if len(r) > mv:
        v = float(r[mv].replace(",", ""))
        if r[mn] == "gpu__time_duration.sum": v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[un], 1)
        recs.setdefault(int(r[idc]), {"k": r[kn].split("(")[0]})[r[mn]] = v
ids = sorted(recs); take = ids[len(ids) // 2:] if which == "second_half" else ids
agg = {}
for i in take:
    r = recs[i]; a = agg.setdefault(r["k"], [0, 0.0, 0.0, 0.0]); a[0] += 1; a[1] += r["gpu__time_duration.sum"]; a[2] += r.get("sm__cycles_active.avg", 0); a[3] += r.get("smsp__inst_executed.sum", 0)
tot = sum(a[1] for a in agg.values()); tota = sum(a[2] for a in agg.values())
out = [f"{len(take)} launches, sum of durations {tot / 1e3:.1f} us (serialised, cold caches), sum of SM-active cycles {tota / 1e3:.0f} kcycles",
       f"{'time us':>10} {'share':>6} {'active kcyc':>12} {'share':>6} {'warp inst M':>12} {'n':>5}  kernel"]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{a[1] / 1e3:10.1f} {100 * a[1] / tot:5.1f}% {a[2] / 1e3:12.1f} {100 * a[2] / max(tota, 1):5.1f}% {a[3] / 1e6:12.2f} {a[0]:5d}  {k}")
txt = "\n".join(out) + "\n"
if len(sys.argv) > 3: open(sys.argv[3], "w").write(txt)
print(txt)
