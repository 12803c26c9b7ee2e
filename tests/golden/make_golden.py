"""Regenerates tests/golden/golden.json from the CPU oracle:  python tests/golden/make_golden.py

For every named case of tests/cases.py it records the sha256 of the inputs (so a drifting
generator is noticed), the full .delta text and sha256 digests of the stage dumps (suffix
array, LCP, anchors, clusters, alignments).  The oracle is the restatement of MUMmer 3.20's
nucmer in oracle/pmn_oracle.c; the reference itself holds no vectors for this path.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np  # noqa: E402
from cases import CASES  # noqa: E402
from oracle import pmn_oracle as O  # noqa: E402


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def record(name):
    ref, qry, kw = CASES[name]()
    r = O.Run(ref, qry, **kw)
    sa, lcp = r.index()
    anc = r.anchors()
    m, off, tag = r.clusters()
    rows, doff, dl = r.alignments()
    return {
        "inputs_sha256": hashlib.sha256(ref + b"\0" + qry).hexdigest(),
        "opts": kw,
        "sa_sha256": digest(sa), "lcp_sha256": digest(lcp),
        "n_anchors": int(len(anc)), "anchors_sha256": digest(anc),
        "n_clusters": int(len(tag)), "clusters_sha256": digest(m, off, tag),
        "n_alignments": int(len(rows)), "alignments_sha256": digest(rows, doff, dl),
        "dp_cells": int(r.dp_cells()),
        "delta": r.delta("ref.fa", "qry.fa").decode(),
    }


if __name__ == "__main__":
    out = {name: record(name) for name in CASES}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(f"{k:32s} anchors {v['n_anchors']:6d} clusters {v['n_clusters']:4d} alignments {v['n_alignments']:4d} delta {len(v['delta']):7d} B")
