"""One-off runs of the other BASELINE.json configurations on one GPU (not bench lines: parity-size cases and
capacity checks).  usage: run_configs.py c4 [n] | c5 [n] | c3 [count]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import lib, synth
what = sys.argv[1]
out = {"config": what}
if what == "c4":
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000_000
    t = time.time(); gs = synth.config_c4(n=n, inv_len=max(1000, n // 100)); out["synth_s"] = round(time.time() - t, 1)
    with lib.Context(0) as ctx:
        t = time.time(); rs, qs = ctx.sequence(synth.fasta(*gs[0])), ctx.sequence(synth.fasta(*gs[1])); out["pack_s"] = round(time.time() - t, 2)
        for rep in range(2):
            t = time.time(); ix = rs.index(); t_ix = time.time() - t
            t = time.time(); res = ix.align(qs, ref_path="g0", qry_path="g1"); t_al = time.time() - t
            st = res.stats
            out[f"rep{rep}"] = {"index_s": round(t_ix, 3), "align_s": round(t_al, 3), "delta_bytes": len(res.delta),
                               **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}}
            res.close(); ix.close()
elif what == "c5":
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 5_000_000
    anc, qs_ = synth.config_c5(n=n)
    with lib.Context(0) as ctx:
        rs = ctx.sequence(synth.fasta(*anc)); ix = rs.index()
        rows = []
        for name, seq in qs_:
            qs = ctx.sequence(synth.fasta(name, seq))
            for rep in range(2):
                res = ix.align(qs, ref_path="anc", qry_path=name); st = res.stats; res.close()
            rows.append({"query": name, "anchors": st["anchors"], "clusters": st["clusters"], "alignments": st["alignments"], "aligned_ref_bases": st["aligned_ref_bases"],
                         "dp_cells": st["dp_cells"], "wave1_cells": st["wave1_cells"], "ms_seed": round(st["ms_seed"], 3), "ms_cluster": round(st["ms_cluster"], 3),
                         "ms_extend": round(st["ms_extend"], 3), "ms_wave1": round(st["ms_wave1"], 3), "ms_stitch": round(st["ms_stitch"], 3), "ms_total": round(st["ms_total"], 3),
                         "wave1_gcups": round(st["wave1_cells"] / max(1e-9, st["ms_wave1"] * 1e-3) / 1e9, 2)})
            qs.close()
        out["sweep"] = rows
elif what == "c3":
    # 57 x 2 Mbp genomes, the 1596 pairs of the job tree (leaves of 7,7,7,7,7,7,7,8 genomes, then cross products up the tree:
    # lib/base/pm_job.ml:43-57,59-77 with max_seqs = 10, paramugsy.ml:29), one scheduler batch per tree level
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 57
    n = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2_000_000
    gs = synth.config_c3(n=n, count=count)
    leaves = [list(range(k * 7, k * 7 + 7)) for k in range(count // 7)]
    leaves[-1] += list(range(len(leaves) * 7, count))
    levels = [[(i, j) for leaf in leaves for a, i in enumerate(leaf) for j in leaf[a + 1:]]]
    groups = leaves
    while len(groups) > 1:
        nxt, pairs = [], []
        for k in range(0, len(groups) - 1, 2):
            pairs += [(i, j) for i in groups[k] for j in groups[k + 1]]
            nxt.append(groups[k] + groups[k + 1])
        if len(groups) % 2: nxt.append(groups[-1])
        levels.append(pairs); groups = nxt
    sched = lib.Scheduler(0, 16); ctx = sched.context(0)
    t = time.time(); seqs = [ctx.sequence(synth.fasta(*g)) for g in gs]; out["pack_s"] = round(time.time() - t, 2)
    names = [g[0] for g in gs]
    out["pairs_per_level"] = [len(l) for l in levels]; out["pairs"] = sum(len(l) for l in levels)
    for rep in range(2):
        t = time.time(); aligned = 0; nal = 0; dbytes = 0
        for pairs in levels:
            for r in sched.align_seqs(seqs, pairs, names=names):
                aligned += r.stats["aligned_ref_bases"]; nal += r.stats["alignments"]; dbytes += len(r.delta); r.close()
        dt = time.time() - t
        out[f"rep{rep}"] = {"wall_s": round(dt, 3), "pairs_per_s": round(out["pairs"] / dt, 1), "aligned_mbp_per_s": round(aligned / 1e6 / dt, 1), "alignments": nal, "delta_bytes": dbytes}
print(json.dumps(out))
