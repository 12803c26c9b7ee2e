"""Per-source-line executed-instruction and stall-sample shares of one kernel from an ncu report.
usage: tools/ncu_lines.py report.ncu-rep kernel_regex cubin_name mangled_prefix [top]"""
import csv, re, subprocess, sys, glob, os, tempfile
rep, kname, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kname], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
h = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[h]; ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
data = [(int(r[ii]), int(r[sm]), r[1]) for r in rows[h + 1:] if len(r) > ii and r[ii].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("paramugsy_b200/_lib/libpmnucmer.so")], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text." + mangled)][0]
cur = None; seq = []
for l in sass[start + 1:]:
    if l.startswith("//---------------------"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2)))
print("sass in cubin", len(seq), "sass in report", len(data))
agg = {}
for k in range(min(len(seq), len(data))):
    a = agg.setdefault(seq[k][0], [0, 0, 0]); a[0] += data[k][0]; a[1] += data[k][1]; a[2] += 1
tot = sum(a[0] for a in agg.values()); tots = max(1, sum(a[1] for a in agg.values()))
src = {}
def getline(f, n):
    if f not in src:
        p = [x for x in glob.glob("paramugsy_b200/csrc/*") + glob.glob("include/*") if x.endswith(f)]
        src[f] = open(p[0]).read().split("\n") if p else []
    return src[f][n - 1].strip()[:120] if src[f] and n <= len(src[f]) else ""
print("total inst", tot, "samples", tots)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/tots:5.1f}% smp {a[2]:4d} sass  {key[0] if key else None}:{key[1] if key else 0}  {getline(*key) if key else ''}")
