"""CPU tests of the oracle itself: every stage is checked against an independent brute-force
definition or a structural property, and the .delta text against the REFERENCE's own parser
(lib/profiles_lib, compiled into oracle/_ref by oracle/build_ref.sh)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from paramugsy_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_pair(n, d, seed, inv=0, inv_len=0):
    g0 = synth.random_genome(n, seed)
    g1 = synth.mutate(g0, d, seed + 1)
    if inv:
        g1 = synth.invert(g1, inv, inv_len, seed + 2)
    return g0, g1


# ------------------------------------------------------------------ index

@pytest.mark.parametrize("n,seed", [(1, 3), (2, 4), (50, 5), (700, 6), (5000, 7)])
def test_sa_is_sorted_permutation_and_lcp_brute(oracle, n, seed):
    g = synth.random_genome(n, seed)
    r = oracle.Run(synth.fasta("a.1", g), synth.fasta("b.1", g))
    sa, lcp = r.index()
    assert sorted(sa.tolist()) == list(range(n))
    s = g  # no X in this text: plain byte order A<C<G<T equals code order, shorter suffix first
    for i in range(1, n):
        assert s[sa[i - 1]:] < s[sa[i]:]
    for i in range(1, min(n, 400)):
        a, b = s[sa[i - 1]:], s[sa[i]:]
        k = 0
        while k < len(a) and k < len(b) and a[k] == b[k]:
            k += 1
        assert lcp[i] == k
    assert n == 0 or lcp[0] == 0


def test_sa_with_separators_and_n_runs(oracle):
    ref = b">r1\nACGTNNNNACGTAC\n>r2\nacgtacgtNAC\n>r3\nTTTT\n"
    r = oracle.Run(ref, b">q\nACGT\n")
    sa, lcp = r.index()
    c = r.ref_codes()
    n = len(c)
    assert n == 14 + 1 + 11 + 1 + 4
    assert sorted(sa.tolist()) == list(range(n))

    def key(i):  # END < acgt < X_p by position
        out = []
        for k in range(i, n):
            if c[k] == 4:
                out.append(4 + k); break
            out.append(int(c[k]))
        return out
    for i in range(1, n):
        assert key(sa[i - 1]) < key(sa[i])
        a, b = key(sa[i - 1]), key(sa[i])
        k = 0
        while k < len(a) and k < len(b) and a[k] == b[k] and a[k] < 4:
            k += 1
        assert lcp[i] == k


# ------------------------------------------------------------------ seeding

@pytest.mark.parametrize("n,d,seed,minmatch", [(600, 0.05, 11, 8), (1500, 0.03, 12, 12), (3000, 0.02, 13, 20),
                                              (2000, 0.10, 14, 10)])
def test_anchors_equal_bruteforce_mums(oracle, n, d, seed, minmatch):
    g0, g1 = small_pair(n, d, seed, inv=1, inv_len=n // 5)
    r = oracle.Run(synth.fasta("a.1", g0), synth.fasta("b.1", g1), minmatch=minmatch)
    a = r.anchors()
    ref = H.codes(g0); q = H.codes(g1)
    want = [(x, y, L, 0) for x, y, L in H.brute_mums(ref, q, minmatch)]
    want += [(x, y, L, 1) for x, y, L in H.brute_mums(ref, H.revcomp_codes(q), minmatch)]
    want.sort(key=lambda t: (t[3], t[1], t[0]))
    assert [tuple(x) for x in a.tolist()] == want
    assert len(want) > 5


def test_anchors_repeats_n_and_multirecord(oracle):
    unit = synth.random_genome(300, 21)
    other = synth.random_genome(400, 22)
    ref = synth.fasta("r.1", unit + b"NNNNN" + other) + synth.fasta("r.2", unit[:150] + synth.random_genome(200, 23))
    qry = synth.fasta("q.1", other[50:350] + b"N" + unit) + synth.fasta("q.2", synth.random_genome(100, 24) + other[:90])
    r = oracle.Run(ref, qry, minmatch=15)
    a = r.anchors()
    rrecs, qrecs = H.parse_fasta(ref), H.parse_fasta(qry)
    rc, _ = H.concat_codes(rrecs)
    want = []
    for k, (_, s) in enumerate(qrecs):
        q = H.codes(s)
        want += [(x, y, L, 2 * k) for x, y, L in H.brute_mums(rc, q, 15)]
        want += [(x, y, L, 2 * k + 1) for x, y, L in H.brute_mums(rc, H.revcomp_codes(q), 15)]
    want.sort(key=lambda t: (t[3], t[1], t[0]))
    assert [tuple(x) for x in a.tolist()] == want
    # unit[:150] occurs twice in the reference, yet the LONGEST match of the query copy (all 300
    # bases) occurs once, so it is reported whole — uniqueness is judged on the longest match
    assert (1, 302, 300, 0) in want


def test_no_anchors_for_unrelated_or_short(oracle):
    r = oracle.Run(synth.fasta("a.1", synth.random_genome(3000, 31)), synth.fasta("b.1", synth.random_genome(3000, 32)))
    assert len(r.anchors()) == 0
    assert r.delta() == b"ref.fa qry.fa\nNUCMER\n"
    r = oracle.Run(b">a\nACGTACGTAC\n", b">b\nACGTACGTAC\n")
    assert len(r.anchors()) == 0  # shorter than minmatch


# ------------------------------------------------------------------ clustering

@pytest.mark.parametrize("n,d,seed", [(20000, 0.02, 41), (30000, 0.05, 42), (30000, 0.10, 43), (20000, 0.15, 44)])
def test_clusters_structure_and_fast_chain_equivalence(oracle, n, d, seed):
    g0, g1 = small_pair(n, d, seed, inv=2, inv_len=1500)
    ra, rb = synth.fasta("a.1", g0), synth.fasta("b.1", g1)
    m0, off0, tag0 = oracle.Run(ra, rb, fast_chain=0).clusters()
    m1, off1, tag1 = oracle.Run(ra, rb, fast_chain=1).clusters()
    assert np.array_equal(m0, m1) and np.array_equal(off0, off1) and np.array_equal(tag0, tag1)
    assert len(tag0) >= 1
    ref = H.codes(g0); qf = H.codes(g1); qr = H.revcomp_codes(qf)
    for k in range(len(tag0)):
        ms = m0[off0[k]:off0[k + 1]]
        assert ms[:, 2].sum() >= 1
        q = qr if tag0[k] & 1 else qf
        for (sa_, sb_, ln) in ms.tolist():
            assert ln >= 1
            assert np.array_equal(ref[sa_ - 1:sa_ - 1 + ln], q[sb_ - 1:sb_ - 1 + ln])  # still an exact match
        for a, b in zip(ms[:-1].tolist(), ms[1:].tolist()):  # trimmed: strictly colinear, no overlap
            assert b[0] >= a[0] + a[2] and b[1] >= a[1] + a[2]


# ------------------------------------------------------------------ extension + delta

@pytest.mark.parametrize("n,d,seed,inv", [(5000, 0.01, 51, 0), (20000, 0.03, 52, 1), (30000, 0.08, 53, 2),
                                          (30000, 0.15, 54, 1), (1000, 0.0, 55, 0)])
def test_alignments_replay_and_error_counts(oracle, n, d, seed, inv):
    g0, g1 = small_pair(n, d, seed, inv=inv, inv_len=2000)
    r = oracle.Run(synth.fasta("a.1", g0), synth.fasta("b.1", g1))
    rows, doff, dl = r.alignments()
    assert len(rows) >= 1
    A = H.codes(g0); Bf = H.codes(g1); Br = H.revcomp_codes(Bf)
    covered = 0
    for k, row in enumerate(rows.tolist()):
        _, _, dirb, sA, eA, sB, eB, err, sim, non = row
        B = Br if dirb else Bf
        errors, cols, first, last = H.walk_alignment(A, B, sA, eA, sB, eB, dl[doff[k]:doff[k + 1]])
        assert errors == err == sim and non == 0
        assert first and last          # alignments begin and end on a matching column
        assert 1 <= sA <= eA <= len(A) and 1 <= sB <= eB <= len(B)
        covered += eA - sA + 1
    if d <= 0.03:
        assert covered >= 0.95 * n
    text = r.delta("/x/a.fa", "/x/b.fa")
    files, ents = H.parse_delta(text)
    assert files == ["/x/a.fa", "/x/b.fa"]
    assert len(ents) == len(rows)
    for (hdr, vals, ds), row, k in zip(ents, rows.tolist(), range(len(rows))):
        assert hdr == ("a.1", "b.1", len(g0), len(g1))
        sB, eB = row[5], row[6]
        if row[2]:
            sB, eB = len(g1) - sB + 1, len(g1) - eB + 1
            assert sB > eB
        assert vals == [row[3], row[4], sB, eB, row[7], row[8], row[9]]
        assert ds == dl[doff[k]:doff[k + 1]].tolist()


def test_identical_sequences_one_full_alignment(oracle):
    g = synth.random_genome(4000, 61)
    t = oracle.nucmer(synth.fasta("a.1", g), synth.fasta("b.1", g))
    assert t == b"ref.fa qry.fa\nNUCMER\n>a.1 b.1 4000 4000\n1 4000 1 4000 0 0 0\n0\n"


def test_reverse_complement_query(oracle):
    g = synth.random_genome(4000, 62)
    rc = bytes(H.revcomp_codes(H.codes(g)).tolist())
    rc = bytes(b"ACGT"[c] for c in rc)
    t = oracle.nucmer(synth.fasta("a.1", g), synth.fasta("b.1", rc))
    assert t == b"ref.fa qry.fa\nNUCMER\n>a.1 b.1 4000 4000\n1 4000 4000 1 0 0 0\n0\n"


def test_multirecord_headers_and_order(oracle):
    u = synth.random_genome(3000, 71); v = synth.random_genome(2500, 72)
    ref = synth.fasta("r.1", u) + synth.fasta("r.2", v)
    qry = synth.fasta("q.1", synth.mutate(v, 0.02, 73)) + synth.fasta("q.2", synth.mutate(u, 0.02, 74))
    _, ents = H.parse_delta(oracle.nucmer(ref, qry))
    hdrs = []
    for hdr, _, _ in ents:
        if not hdrs or hdrs[-1] != hdr[:2]:
            hdrs.append(hdr[:2])
    assert hdrs == [("r.2", "q.1"), ("r.1", "q.2")]


def test_options_change_results(oracle):
    g0, g1 = small_pair(20000, 0.06, 81)
    ra, rb = synth.fasta("a.1", g0), synth.fasta("b.1", g1)
    base = oracle.Run(ra, rb).anchors()
    assert len(oracle.Run(ra, rb, minmatch=30).anchors()) < len(base)
    assert len(oracle.Run(ra, rb, do_reverse=0).anchors()) <= len(base)
    n_ext = len(oracle.Run(ra, rb).alignments()[0])
    n_noext = len(oracle.Run(ra, rb, do_extend=0).alignments()[0])
    assert n_noext >= n_ext
    with pytest.raises(RuntimeError):
        oracle.Run(ra, rb, do_optimize=0).alignments()


def test_bad_fasta_is_an_error(oracle):
    with pytest.raises(ValueError):
        oracle.Run(b"ACGT\n", b">q\nACGT\n")
    with pytest.raises(ValueError):
        oracle.Run(b"", b">q\nACGT\n")


# ------------------------------------------------------------------ format contract vs the reference parser

REF_RT = os.path.join(ROOT, "oracle", "_ref", "ref_delta_roundtrip")
REF_PRINT = os.path.join(ROOT, "oracle", "_ref", "m_delta_stream_test")


@pytest.fixture(scope="module")
def ref_tools():
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["sh", os.path.join(ROOT, "oracle", "build_ref.sh")])
    if not os.path.exists(REF_RT):
        pytest.skip("oracle/_ref not built and /root/reference absent: format parity vs reference parser UNVERIFIED here")
    return REF_RT, REF_PRINT


def test_reference_known_answer_m_delta_cc_43_49(ref_tools, tmp_path):
    """The only numeric known-answer the reference holds for this format:
    lib/profiles_lib/m_delta.cc:43-49."""
    p = tmp_path / "k.delta"
    p.write_bytes(b"r.fa q.fa\nNUCMER\n>a b 5000 5000\n1 2000 1 2000 8 8 0\n106\n-6\n1797\n-9\n-9\n-1\n7\n1\n0\n")
    out = subprocess.check_output([ref_tools[1], str(p)]).decode()
    ref_gaps = out.split("ref_gaps\n")[1].split("query_range")[0].split()
    qry_gaps = out.split("query_gaps\n")[1].split("\n106")[0].split()
    assert ref_gaps == ["(112,", "112)", "(1918,", "1918)", "(1927,", "1928)"]
    assert qry_gaps == ["(106,", "106)", "(1909,", "1909)", "(1935,", "1936)"]


@pytest.mark.parametrize("n,d,seed,inv", [(20000, 0.03, 91, 1), (30000, 0.10, 92, 2)])
def test_delta_parses_and_reencodes_in_reference_reader(oracle, ref_tools, tmp_path, n, d, seed, inv):
    g0, g1 = small_pair(n, d, seed, inv=inv, inv_len=2500)
    text = oracle.nucmer(synth.fasta("a.1", g0), synth.fasta("b.1", g1), "/d/a.fa", "/d/b.fa")
    p = tmp_path / "o.delta"; p.write_bytes(text)
    out = subprocess.check_output([ref_tools[0], str(p)], stderr=subprocess.DEVNULL)
    # the reference writer prints "1 2 3" for the error counts and (quirk, m_delta.cc:79) the
    # second path for both file names; everything else must round-trip byte for byte
    mine = []
    for k, line in enumerate(text.decode().split("\n")):
        t = line.split(" ")
        if k == 0:
            line = t[-1]
        elif len(t) == 7:
            line = " ".join(t[:4] + ["1", "2", "3"])
        mine.append(line)
    assert out.decode() == "\n".join(mine)
