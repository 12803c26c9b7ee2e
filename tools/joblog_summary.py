"""Summarise a PMN_JOBLOG file: per kernel, calls / cells / cycles, the slowest calls."""
import sys, collections
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1]) if l.strip() and not l.startswith("#")]
for kid, name in ((1, "wave1"), (3, "wave2"), (2, "stitch")):
    r = [x for x in rows if x[6] == kid]
    if not r: continue
    cyc = sum(x[5] for x in r); cells = sum(x[4] for x in r); diags = sum(x[3] for x in r)
    print(f"{name}: calls {len(r)} cells {cells} diagonals {diags} cycles {cyc} cycles/diag {cyc / max(1, diags):.1f} cells/diag {cells / max(1, diags):.1f} max cycles {max(x[5] for x in r)}")
    by = collections.Counter(); cy = collections.Counter()
    for x in r: by[(x[0], x[7])] += 1; cy[(x[0], x[7])] += x[5]
    for k in sorted(by): print(f"   m_o {k[0]:#x} path {k[1]}: {by[k]} calls, {cy[k]} cycles")
    print("   slowest:", *[f"(m_o={x[0]:#x} N={x[1]} M={x[2]} d={x[3]} cells={x[4]} cyc={x[5]} path={x[7]})" for x in sorted(r, key=lambda x: -x[5])[:12]], sep="\n     ")
    h = collections.Counter()
    for x in r: h[min(20, x[5].bit_length())] += 1
    print("   log2(cycles) histogram:", dict(sorted(h.items())))
