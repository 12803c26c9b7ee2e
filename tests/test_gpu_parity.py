"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs and against the committed golden vectors.  Bit-exact at every stage:
suffix array, LCP, anchors, clusters, alignment rows, delta lists and the .delta text."""
import json
import os

import numpy as np
import pytest

import helpers as H
from cases import CASES
from paramugsy_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.fixture(scope="module")
def ctx():
    from paramugsy_b200 import lib
    c = lib.Context(0)      # raises without a B200: there is no fallback
    yield c
    c.close()


def run_gpu(ctx, ref, qry, **kw):
    rs, qs = ctx.sequence(ref), ctx.sequence(qry)
    ix = rs.index()
    res = ix.align(qs, ref_path="ref.fa", qry_path="qry.fa", keep_stages=1, **kw)
    out = dict(sa_lcp=ix.suffix_array(), anchors=res.anchors(), clusters=res.clusters(), alignments=res.alignments(),
               delta=res.delta, stats=res.stats)
    res.close(); ix.close(); qs.close(); rs.close()
    return out


def assert_same_as_oracle(oracle, g, ref, qry, **kw):
    r = oracle.Run(ref, qry, fast_chain=1, **kw)
    sa, lcp = r.index()
    assert np.array_equal(g["sa_lcp"][0], sa), "suffix array differs"
    assert np.array_equal(g["sa_lcp"][1], lcp), "LCP differs"
    assert np.array_equal(g["anchors"], r.anchors()), "anchors differ"
    for a, b, what in zip(g["clusters"], r.clusters(), ("cluster matches", "cluster offsets", "cluster tags")):
        assert np.array_equal(a, b), what + " differ"
    for a, b, what in zip(g["alignments"], r.alignments(), ("alignment rows", "delta offsets", "deltas")):
        assert np.array_equal(a, b), what + " differ"
    assert g["delta"] == r.delta("ref.fa", "qry.fa"), ".delta text differs"
    # dp_cells is NOT a parity quantity: wave 1 aligns every match -> next match gap up front, including
    # those of clusters that extendClusters later skips as shadowed (tools/cells_check.py lists both counts)


@pytest.mark.parametrize("name", sorted(CASES))
def test_every_stage_equals_oracle_and_golden(ctx, oracle, name):
    ref, qry, kw = CASES[name]()
    g = run_gpu(ctx, ref, qry, **kw)
    assert g["delta"].decode() == GOLD[name]["delta"], "differs from the committed golden .delta"
    assert_same_as_oracle(oracle, g, ref, qry, **kw)
    assert g["stats"]["kernel_launches"] > 0


def test_config1_two_1mbp_genomes(ctx, oracle):
    """BASELINE.json configs[0]: two synthetic 1 Mbp genomes, 1 % SNP/indel."""
    gs = synth.config_c1()
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    g = run_gpu(ctx, ref, qry)
    assert_same_as_oracle(oracle, g, ref, qry)
    assert g["stats"]["alignments"] >= 1 and g["stats"]["aligned_ref_bases"] > 990_000


def test_divergence_sweep_500k(ctx, oracle):
    """configs[4] at reduced length: the extension-bound end of the sweep must stay bit-exact."""
    anc = synth.random_genome(500_000, 5000)
    ref = synth.fasta("anc.1", anc)
    for d in (0.05, 0.10, 0.15):
        qry = synth.fasta("q.1", synth.mutate(anc, d, 5000 + int(1000 * d)))
        g = run_gpu(ctx, ref, qry)
        assert_same_as_oracle(oracle, g, ref, qry)


def test_index_reuse_across_queries_and_repeatability(ctx, oracle):
    """One index, several queries (how a reference genome serves its n-1-i pairs,
    lib/base/pm_job.ml:43-51); results do not depend on what ran before."""
    gs = synth.config_c2(n=150_000, count=4, inv_len=4_000)
    rs = ctx.sequence(synth.fasta(*gs[0]))
    ix = rs.index()
    first = {}
    for rnd in range(2):
        for name, seq in gs[1:]:
            qs = ctx.sequence(synth.fasta(name, seq))
            res = ix.align(qs, ref_path="r", qry_path="q")
            if rnd == 0:
                first[name] = res.delta
                assert res.delta == oracle.nucmer(synth.fasta(*gs[0]), synth.fasta(name, seq), "r", "q", fast_chain=1)
            else:
                assert res.delta == first[name]
            res.close(); qs.close()
    ix.close(); rs.close()


def test_full_size_properties_5mbp_pair(ctx):
    """Size-independent properties at BASELINE.json's full pair size (8 x 5 Mbp config), where
    the oracle would take minutes: every alignment replays exactly over the sequences, error
    counts match a recount, rows are sane, and the run is deterministic."""
    gs = synth.config_c2(count=2)
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    g = run_gpu(ctx, ref, qry)
    rows, doff, dl = g["alignments"]
    A = H.codes(gs[0][1]); Bf = H.codes(gs[1][1]); Br = H.revcomp_codes(Bf)
    assert len(rows) >= 3
    covered = 0
    for k, row in enumerate(rows.tolist()):
        _, _, dirb, sA, eA, sB, eB, err, sim, non = row
        assert 1 <= sA <= eA <= len(A) and 1 <= sB <= eB <= len(Bf)
        if eA - sA < 200_000:       # the pure-Python replay is slow; recount the shorter ones fully
            errors, cols, first, last = H.walk_alignment(A, Br if dirb else Bf, sA, eA, sB, eB, dl[doff[k]:doff[k + 1]])
            assert errors == err == sim and non == 0 and first and last
        covered += eA - sA + 1
    assert covered > 0.97 * len(A)
    # anchors are exact, unique-in-reference matches (checked on a sample)
    anc = g["anchors"]
    assert len(anc) > 50_000
    for r_, q_, ln, tag in anc[:: max(1, len(anc) // 500)].tolist():
        B = Br if tag & 1 else Bf
        assert np.array_equal(A[r_ - 1:r_ - 1 + ln], B[q_ - 1:q_ - 1 + ln]) and ln >= 20
    assert run_gpu(ctx, ref, qry)["delta"] == g["delta"]
    files, ents = H.parse_delta(g["delta"])
    assert len(ents) == len(rows)


def test_file_level_entry_points(ctx, oracle, tmp_path):
    """pmn_align_pair / pmn_align_batch: what the OCaml stub binds."""
    import ctypes as C
    from paramugsy_b200 import lib
    gs = synth.config_c2(n=60_000, count=3, inv_len=2_000)
    paths = []
    for name, seq in gs:
        p = tmp_path / (name + ".fa"); p.write_bytes(synth.fasta(name, seq)); paths.append(str(p))
    pairs = [(paths[0], paths[1]), (paths[0], paths[2]), (paths[1], paths[2])]
    outs = [str(tmp_path / f"{os.path.basename(a)}-{os.path.basename(b)}.delta") for a, b in pairs]
    arr = lambda xs: (C.c_char_p * len(xs))(*[x.encode() for x in xs])
    rc = lib.lib().pmn_align_batch(ctx.h, len(pairs), arr([a for a, _ in pairs]), arr([b for _, b in pairs]), arr(outs), None)
    assert rc == 0, lib.lib().pmn_last_error(None)
    for (a, b), o in zip(pairs, outs):
        want = oracle.nucmer(open(a, "rb").read(), open(b, "rb").read(), a, b, fast_chain=1)
        assert open(o, "rb").read() == want
        assert not [f for f in os.listdir(tmp_path) if ".tmp." in f]
    rc = lib.lib().pmn_align_pair(ctx.h, b"/nonexistent/ref.fa", paths[1].encode(), None, outs[0].encode())
    assert rc == -4 and b"cannot open" in lib.lib().pmn_last_error(None)


def test_error_paths(ctx):
    from paramugsy_b200 import lib
    with pytest.raises(lib.PmnError):
        ctx.sequence(b"ACGT\n")                      # data before a header
    with pytest.raises(lib.PmnError):
        ctx.sequence(b"")
    rs = ctx.sequence(b">a\nACGTACGTACGTACGTACGTAAAACCCCGGGG\n")
    ix = rs.index()
    with pytest.raises(lib.PmnError):
        ix.align(rs, do_optimize=0)                  # --nooptimize is not implemented
    with pytest.raises(lib.PmnError):
        ix.align(rs, minmatch=0)
    ix.close(); rs.close()


def test_scheduler_results_do_not_depend_on_workers(oracle, tmp_path):
    """pmn_sched: the in-process form of run_nucmers (lib/base/job_processor.ml:128-154).  Several
    pairs in flight on one GPU, shared indexes; every .delta equals the oracle's whatever the
    number of workers, from host FASTA bytes, from resident genomes and from files."""
    from paramugsy_b200 import lib
    gs = synth.config_c2(n=80_000, count=5, inv_len=2_500)
    fastas = [synth.fasta(n, s) for n, s in gs]
    names = [n for n, _ in gs]
    pairs = [(i, j) for i in range(5) for j in range(i + 1, 5)] + [(2, 2)]
    want = [oracle.nucmer(fastas[i], fastas[j], names[i], names[j], fast_chain=1) for i, j in pairs]
    for workers in (1, 3):
        with lib.Scheduler(0, workers) as s:
            res = s.align_fasta(fastas, pairs, names=names)
            assert [r.delta for r in res] == want
            assert s.counters()["pairs"] == len(pairs) and s.counters()["launches"] > 0
            c = s.context(0)
            seqs = [c.sequence(f) for f in fastas]
            res = s.align_seqs(seqs, pairs, names=names)
            assert [r.delta for r in res] == want
            for q in seqs:
                q.close()
            paths = []
            for n, f in zip(names, fastas):
                p = tmp_path / f"{workers}_{n}.fa"; p.write_bytes(f); paths.append(str(p))
            outs = [str(tmp_path / f"{workers}_{i}_{j}.delta") for i, j in pairs]
            s.align_files([paths[i] for i, _ in pairs], [paths[j] for _, j in pairs], outs)
            for (i, j), o in zip(pairs, outs):
                assert open(o, "rb").read() == oracle.nucmer(fastas[i], fastas[j], paths[i], paths[j], fast_chain=1)
            with pytest.raises(lib.PmnError):
                s.align_fasta(fastas[:2] + [b"not fasta"], [(0, 1), (0, 2)])
            assert [r.delta for r in s.align_fasta(fastas, pairs[:3], names=names)] == want[:3]     # usable after an error


def test_sharded_seeding_and_replicated_index(ctx, oracle):
    """SURVEY.md §8e on one GPU: (1) the anchors of query-position parts, concatenated in part order,
    are the anchor list of the undivided run, and the alignment continued from them is the same
    .delta for any number of parts; (2) an index image copied into an empty index of another context
    (what the NCCL broadcast does between GPUs) serves the same results."""
    import torch
    from paramugsy_b200 import lib
    gs = synth.config_c2(n=300_000, count=2, inv_len=6_000)
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1]) + synth.fasta("extra.1", synth.random_genome(700, 5))
    rs, qs = ctx.sequence(ref), ctx.sequence(qry)
    ix = rs.index()
    whole = ix.align(qs, ref_path="r", qry_path="q", keep_stages=1)
    want_anchors = torch.from_numpy(whole.anchors())
    want = whole.delta
    assert want == oracle.nucmer(ref, qry, "r", "q", fast_chain=1)
    for parts in (1, 2, 3, 8):
        got = torch.cat([ix.seed_part_tensor(qs, k, parts).cpu() for k in range(parts)])
        assert torch.equal(got, want_anchors), f"{parts} parts"
        res = ix.align_anchors(qs, got.cuda(), ref_path="r", qry_path="q")
        assert res.delta == want
        res.close()
    empty = ix.align_anchors(qs, torch.empty((0, 4), dtype=torch.int32, device="cuda"), ref_path="r", qry_path="q")
    assert empty.delta == b"r q\nNUCMER\n"
    # replicated index
    with lib.Context(0) as other:
        rs2, qs2 = other.sequence(ref), other.sequence(qry)
        ix2 = rs2.index(empty=True)
        with pytest.raises(lib.PmnError):
            ix2.adopt()                                   # nothing received yet
        assert ix2.image()[1] == ix.image()[1] == lib.index_image_bytes(rs.bases)
        ix2.image_tensor().copy_(ix.image_tensor())
        torch.cuda.synchronize()
        ix2.adopt()
        r2 = ix2.align(qs2, ref_path="r", qry_path="q")
        assert r2.delta == want
        r2.close(); ix2.close(); qs2.close(); rs2.close()
    whole.close(); ix.close(); qs.close(); rs.close()
