"""Regenerates tests/golden/golden_post.json from the CPU oracle:  python tests/golden/make_golden_post.py

For every case of tests/post_cases.py: sha256 of `delta-filter -1`, `delta-filter -m` and `delta2maf` of the
oracle's .delta, and the alignment counts (input, -1, -m).  The programs are external to the reference
(MUMmer 3.20 / Mugsy) and it holds no vectors for them; rules in ORACLE_SPEC.md §8-§9.
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from post_cases import POST_CASES  # noqa: E402
from oracle import pmn_oracle as O  # noqa: E402


def nal(t):
    return sum(1 for l in t.split(b"\n") if l.count(b" ") == 6 and not l.startswith(b">"))


def record(name):
    ref, qry, kw = POST_CASES[name]()
    d = O.nucmer(ref, qry, "ref.fa", "qry.fa", fast_chain=1, **kw)
    one, many = O.delta_filter(d, 1), O.delta_filter(d, 2)
    return {"delta_sha256": hashlib.sha256(d).hexdigest(), "filter1_sha256": hashlib.sha256(one).hexdigest(), "filterm_sha256": hashlib.sha256(many).hexdigest(),
            "maf_sha256": hashlib.sha256(O.delta2maf(d, ref, qry)).hexdigest(), "n_alignments": [nal(d), nal(one), nal(many)]}


if __name__ == "__main__":
    out = {name: record(name) for name in POST_CASES}
    with open(os.path.join(HERE, "golden_post.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in sorted(out.items()):
        print(k, v["n_alignments"])
