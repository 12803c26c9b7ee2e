"""The oracle against the committed golden vectors (CPU).  The vectors were produced by
tests/golden/make_golden.py; a mismatch means the oracle or the genome generator drifted."""
import hashlib
import json
import os

import numpy as np
import pytest

from cases import CASES

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_golden_covers_all_cases():
    assert sorted(GOLD) == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(oracle, name):
    ref, qry, kw = CASES[name]()
    g = GOLD[name]
    assert hashlib.sha256(ref + b"\0" + qry).hexdigest() == g["inputs_sha256"], "synthetic inputs drifted"
    assert kw == g["opts"]
    r = oracle.Run(ref, qry, **kw)
    sa, lcp = r.index()
    assert digest(sa) == g["sa_sha256"] and digest(lcp) == g["lcp_sha256"]
    anc = r.anchors()
    assert len(anc) == g["n_anchors"] and digest(anc) == g["anchors_sha256"]
    m, off, tag = r.clusters()
    assert len(tag) == g["n_clusters"] and digest(m, off, tag) == g["clusters_sha256"]
    rows, doff, dl = r.alignments()
    assert len(rows) == g["n_alignments"] and digest(rows, doff, dl) == g["alignments_sha256"]
    assert r.delta("ref.fa", "qry.fa").decode() == g["delta"]
    # the literal O(m^2) chain DP and the pruned one agree
    assert oracle.Run(ref, qry, fast_chain=1, **kw).delta("ref.fa", "qry.fa").decode() == g["delta"]
