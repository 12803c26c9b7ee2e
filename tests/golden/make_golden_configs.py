"""Regenerates tests/golden/golden_configs.json from the CPU oracle, at the FULL sizes BASELINE.json names:

    python tests/golden/make_golden_configs.py [c1] [c2] [c3] [c5] [c4] [--jobs N]

  C1  the 1 Mbp pair
  C2  all 28 pairs of the 8 x 5 Mbp all-vs-all (earlier genome = reference, lib/base/pm_job.ml:43-51)
  C3  24 pairs of the 57 x 2 Mbp job tree: 12 leaf pairs (both genomes inside one leaf of the max_seqs = 10 tree,
      lib/base/pm_job.ml:59-77) and 12 cross pairs (left x right of a merge node, pm_job.ml:53-57) — fixed list below
  C5  the 8 divergence points 1 % ... 15 % on 5 Mbp
  C4  the 100 Mbp pair (one oracle run: tens of minutes on one core and ~3 GB)

For every pair it records the sha256 of the inputs (a drifting generator is noticed), sha256 + length of the .delta
text, of `delta-filter -1` of it and of the MAF of the filtered delta, and the alignment / anchor / cluster counts.
Only digests are committed: the texts are 0.3-20 MB each.  Line 1 of every .delta is "<ref name>.fa <qry name>.fa".
The oracle is the restatement of MUMmer 3.20's nucmer in oracle/pmn_oracle.c (PARITY UNPINNED vs MUMmer itself,
SURVEY.md §8c); the configs without an argument are merged into the existing file, so C4 can be run alone.
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from paramugsy_b200 import synth  # noqa: E402

OUT = os.path.join(HERE, "golden_configs.json")

# C3 sample: leaves of the job tree hold genomes 0-6, 7-13, ..., 42-48, 49-56
C3_LEAF = [(0, 1), (0, 6), (3, 5), (7, 8), (9, 13), (14, 20), (22, 27), (28, 34), (35, 41), (42, 48), (49, 50), (51, 56)]
C3_CROSS = [(0, 7), (6, 13), (2, 20), (14, 27), (5, 30), (10, 40), (28, 48), (0, 56), (13, 49), (21, 56), (33, 52), (27, 28)]


def sha(b):
    return hashlib.sha256(b).hexdigest()


def _pair_record(args):
    """Runs in a worker process; genomes are regenerated there from the config so nothing large is pickled."""
    cfg, key, ia, ib = args
    from oracle import pmn_oracle as O
    gs = genomes_of(cfg)
    (na, sa), (nb, sb) = gs[ia], gs[ib]
    ref, qry = synth.fasta(na, sa), synth.fasta(nb, sb)
    t0 = time.time()
    r = O.Run(ref, qry, fast_chain=1)
    n_anchors = int(len(r.anchors()))
    n_clusters = int(len(r.clusters()[2]))
    delta = r.delta(na + ".fa", nb + ".fa")
    n_align = int(len(r.alignments()[0]))
    cells = int(r.dp_cells())
    r.close()
    filt = O.delta_filter(delta, 1)
    maf = O.delta2maf(filt, ref, qry)
    rec = {"ref": na, "qry": nb, "inputs_sha256": sha(ref + b"\0" + qry),
           "n_anchors": n_anchors, "n_clusters": n_clusters, "n_alignments": n_align, "dp_cells": cells,
           "delta_bytes": len(delta), "delta_sha256": sha(delta),
           "filtered_bytes": len(filt), "filtered_sha256": sha(filt),
           "maf_bytes": len(maf), "maf_sha256": sha(maf),
           "oracle_seconds": round(time.time() - t0, 2)}
    return cfg, key, rec


_GENOMES = {}


def genomes_of(cfg):
    """Full-size genome list of a config (cached per process)."""
    if cfg not in _GENOMES:
        if cfg == "c1":
            _GENOMES[cfg] = synth.config_c1()
        elif cfg == "c2":
            _GENOMES[cfg] = synth.config_c2()
        elif cfg == "c3":
            _GENOMES[cfg] = synth.config_c3()
        elif cfg == "c4":
            _GENOMES[cfg] = synth.config_c4()
        elif cfg == "c5":
            anc, qs = synth.config_c5()
            _GENOMES[cfg] = [anc] + qs
        else:
            raise KeyError(cfg)
    return _GENOMES[cfg]


def jobs_of(cfg):
    if cfg == "c1":
        return [(cfg, "g0.1-g1.1", 0, 1)]
    if cfg == "c2":
        return [(cfg, f"g{i}.1-g{j}.1", i, j) for i in range(8) for j in range(i + 1, 8)]
    if cfg == "c3":
        return [(cfg, f"s{i}.1-s{j}.1", i, j) for i, j in C3_LEAF + C3_CROSS]
    if cfg == "c4":
        return [(cfg, "c0.1-c1.1", 0, 1)]
    if cfg == "c5":
        names = [n for n, _ in genomes_of("c5")]
        return [(cfg, f"{names[0]}-{names[k]}", 0, k) for k in range(1, len(names))]
    raise KeyError(cfg)


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    jobs = 8
    for k, a in enumerate(sys.argv):
        if a == "--jobs":
            jobs = int(sys.argv[k + 1]); argv = [x for x in argv if x != sys.argv[k + 1]]
    cfgs = argv or ["c1", "c2", "c3", "c5"]
    out = json.load(open(OUT)) if os.path.exists(OUT) else {}
    work = [j for c in cfgs for j in jobs_of(c)]
    for c in cfgs:
        out[c] = {}
    with mp.get_context("fork").Pool(min(jobs, len(work))) as pool:
        for cfg, key, rec in pool.imap_unordered(_pair_record, work):
            out[cfg][key] = rec
            print(f"{cfg} {key:20s} anchors {rec['n_anchors']:8d} alignments {rec['n_alignments']:6d} delta {rec['delta_bytes']:9d} B  {rec['oracle_seconds']:.1f} s", flush=True)
    out["_meta"] = {"made_by": "tests/golden/make_golden_configs.py", "oracle": "oracle/pmn_oracle.c (fast_chain=1)",
                    "c3_leaf_pairs": C3_LEAF, "c3_cross_pairs": C3_CROSS, "line1": "<ref name>.fa <qry name>.fa"}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
