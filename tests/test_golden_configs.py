"""tests/golden/golden_configs.json: the oracle's digests at the full BASELINE sizes (made by make_golden_configs.py).
CPU side: the file is complete, and the oracle still reproduces it on the cases that take seconds (C1 and one C3 pair);
the GPU side is tests/test_gpu_configs.py."""
import hashlib
import json
import os

from paramugsy_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_configs.json")))


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_every_config_is_there_at_full_size():
    assert len(GOLD["c1"]) == 1 and len(GOLD["c2"]) == 28 and len(GOLD["c3"]) >= 20 and len(GOLD["c5"]) == 8
    assert {(r["ref"], r["qry"]) for r in GOLD["c2"].values()} == {(f"g{i}.1", f"g{j}.1") for i in range(8) for j in range(i + 1, 8)}
    for cfg in ("c1", "c2", "c3", "c5") + (("c4",) if "c4" in GOLD else ()):
        for key, r in GOLD[cfg].items():
            assert len(r["delta_sha256"]) == 64 and r["delta_bytes"] > 100 and r["n_alignments"] >= 1, (cfg, key)
    if "c4" in GOLD:
        assert GOLD["c4"]["c0.1-c1.1"]["n_anchors"] > 500_000


def _check(oracle, cfg, key, a, b):
    g = GOLD[cfg][key]
    ref, qry = synth.fasta(*a), synth.fasta(*b)
    assert sha(ref + b"\0" + qry) == g["inputs_sha256"], "the synthetic genomes drifted"
    d = oracle.nucmer(ref, qry, a[0] + ".fa", b[0] + ".fa", fast_chain=1)
    assert sha(d) == g["delta_sha256"] and len(d) == g["delta_bytes"]
    f = oracle.delta_filter(d, 1)
    assert sha(f) == g["filtered_sha256"]
    assert sha(oracle.delta2maf(f, ref, qry)) == g["maf_sha256"]


def test_oracle_reproduces_c1(oracle):
    gs = synth.config_c1()
    _check(oracle, "c1", "g0.1-g1.1", gs[0], gs[1])


def test_oracle_reproduces_a_c3_pair_with_the_unpruned_chain_dp(oracle):
    """The digests were made with fast_chain=1 (pruned chain DP); the literal O(m^2) one gives the same bytes."""
    gs = synth.config_c3(count=2)
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    g = GOLD["c3"]["s0.1-s1.1"]
    assert sha(oracle.nucmer(ref, qry, "s0.1.fa", "s1.1.fa", fast_chain=0)) == g["delta_sha256"]
