"""GPU parity at the FULL sizes BASELINE.json names (-m gpu): the CUDA path through the C ABI against the sha256
digests of the oracle's output committed in tests/golden/golden_configs.json (made by
tests/golden/make_golden_configs.py in the CPU container; the oracle needs 20 s per 5 Mbp pair and tens of minutes
for the 100 Mbp pair, the digests travel instead).  Bit-exact: the .delta text of every pair, and for the
all-vs-all configs also `delta-filter -1` of it and the MAF of the filtered delta (what one reference worker
leaves behind, lib/nucmer/mugsy_nucmer.ml:127-131).

  C1  two 1 Mbp genomes                         1 pair
  C2  8 x 5 Mbp all-vs-all                      all 28 pairs, through the batch scheduler
  C3  57 x 2 Mbp job tree                       the 24 pairs (12 leaf, 12 cross) listed in the golden file
  C4  one 100 Mbp pair                          undivided and with the query positions cut into 2, 4 and 8 parts
                                                (the partition of SURVEY.md §8e: parts' anchor lists concatenated)
  C5  5 Mbp pairs at 1 % ... 15 % divergence    all 8 points
"""
import hashlib
import json
import os

import pytest

from paramugsy_b200 import synth

pytestmark = pytest.mark.gpu

GOLD_PATH = os.path.join(os.path.dirname(__file__), "golden", "golden_configs.json")
GOLD = json.load(open(GOLD_PATH)) if os.path.exists(GOLD_PATH) else {}


def sha(b):
    return hashlib.sha256(b).hexdigest()


@pytest.fixture(scope="module")
def sched():
    from paramugsy_b200 import lib
    s = lib.Scheduler(0, 8)      # raises without a B200: there is no fallback
    yield s
    s.close()


def check_pairs(sched, cfg, genomes, jobs):
    """jobs: [(key, i, j)] over `genomes` [(name, seq)]; one scheduler call with the post-steps on."""
    gold = GOLD[cfg]
    used = sorted({g for _, i, j in jobs for g in (i, j)})
    local = {g: k for k, g in enumerate(used)}
    fastas = [synth.fasta(*genomes[g]) for g in used]
    names = [genomes[g][0] + ".fa" for g in used]
    for key, i, j in jobs:
        assert sha(fastas[local[i]] + b"\0" + fastas[local[j]]) == gold[key]["inputs_sha256"], f"{cfg} {key}: the synthetic genomes drifted"
    res = sched.align_fasta(fastas, [(local[i], local[j]) for _, i, j in jobs], names=names, post=1)
    bad = []
    for (key, i, j), r in zip(jobs, res):
        g = gold[key]
        d, f, m = r.delta, r.filtered, r.maf
        st = r.stats
        r.close()
        for what, got, want, n in (("delta", sha(d), g["delta_sha256"], len(d)), ("filtered", sha(f), g["filtered_sha256"], len(f)), ("maf", sha(m), g["maf_sha256"], len(m))):
            if got != want:
                bad.append(f"{key}: {what} differs ({n} bytes, oracle {g[what + '_bytes']}; anchors {st['anchors']} vs {g['n_anchors']}, "
                           f"clusters {st['clusters']} vs {g['n_clusters']}, alignments {st['alignments']} vs {g['n_alignments']})")
    assert not bad, f"{cfg}: {len(bad)} texts differ from the oracle's: " + "; ".join(bad[:6])


def test_c1_pair_1mbp(sched):
    check_pairs(sched, "c1", synth.config_c1(), [("g0.1-g1.1", 0, 1)])


def test_c2_all_28_pairs_5mbp(sched):
    gs = synth.config_c2()
    jobs = [(f"g{i}.1-g{j}.1", i, j) for i in range(8) for j in range(i + 1, 8)]
    assert len(jobs) == 28 and set(GOLD["c2"]) == {k for k, _, _ in jobs}
    check_pairs(sched, "c2", gs, jobs)


def test_c3_sample_of_the_job_tree_2mbp(sched):
    meta = GOLD["_meta"]
    sample = [tuple(p) for p in meta["c3_leaf_pairs"] + meta["c3_cross_pairs"]]
    assert len(sample) >= 20
    gs = synth.config_c3()
    check_pairs(sched, "c3", gs, [(f"s{i}.1-s{j}.1", i, j) for i, j in sample])


def test_c5_divergence_sweep_5mbp(sched):
    anc, qs = synth.config_c5()
    gs = [anc] + qs
    assert len(qs) == 8
    check_pairs(sched, "c5", gs, [(f"{anc[0]}-{q[0]}", 0, k + 1) for k, q in enumerate(qs)])


@pytest.mark.skipif("c4" not in GOLD, reason="golden_configs.json holds no C4 digest yet (make_golden_configs.py c4)")
def test_c4_100mbp_pair_and_its_query_partitions():
    """The undivided run and the runs seeded in 2, 4 and 8 query-position parts must all give the oracle's bytes."""
    import torch
    from paramugsy_b200 import lib
    g = GOLD["c4"]["c0.1-c1.1"]
    gs = synth.config_c4()
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    assert sha(ref + b"\0" + qry) == g["inputs_sha256"], "c4: the synthetic genomes drifted"
    with lib.Context(0) as ctx:
        rs, qs = ctx.sequence(ref), ctx.sequence(qry)
        del ref, qry
        ix = rs.index()
        res = ix.align(qs, ref_path="c0.1.fa", qry_path="c1.1.fa")
        d, st = res.delta, res.stats
        res.close()
        assert st["anchors"] == g["n_anchors"] and st["alignments"] == g["n_alignments"], (st, g)
        assert sha(d) == g["delta_sha256"], "c4: .delta differs from the oracle's"
        for parts in (2, 4, 8):
            anchors = torch.cat([ix.seed_part_tensor(qs, k, parts) for k in range(parts)])
            assert anchors.shape[0] == g["n_anchors"]
            res = ix.align_anchors(qs, anchors, ref_path="c0.1.fa", qry_path="c1.1.fa")
            dp = res.delta
            res.close()
            assert sha(dp) == g["delta_sha256"], f"c4: .delta differs when the query is seeded in {parts} parts"
        ix.close(); qs.close(); rs.close()
