"""The two post-steps of lib/nucmer/mugsy_nucmer.ml (delta-filter :102-105, delta2maf :118-124) in the CPU
oracle: size-independent properties, known answers, and the grammars the reference's own readers accept."""
import hashlib
import json
import os
import re
import subprocess

import pytest

import helpers as H
from post_cases import POST_CASES

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_post.json")))
REF_TOOL = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_delta_roundtrip")


def alignments(delta: bytes):
    """[(block header, 7 ints, [deltas])] of a .delta text."""
    out, hdr, cur = [], None, None
    for line in delta.split(b"\n")[2:]:
        if not line:
            continue
        if line.startswith(b">"):
            hdr = line
        elif cur is None:
            cur = (hdr, tuple(int(x) for x in line.split()), [])
            assert len(cur[1]) == 7
        elif line == b"0":
            out.append(cur); cur = None
        else:
            cur[2].append(int(line))
    assert cur is None
    return out


@pytest.fixture(scope="module")
def deltas(oracle):
    d = {}
    for name, f in POST_CASES.items():
        ref, qry, kw = f()
        d[name] = (ref, qry, oracle.nucmer(ref, qry, "ref.fa", "qry.fa", fast_chain=1, **kw))
    return d


def test_golden_post_covers_all_cases():
    assert sorted(GOLD) == sorted(POST_CASES)


@pytest.mark.parametrize("name", sorted(POST_CASES))
def test_filter_properties_and_golden(oracle, deltas, name):
    ref, qry, d = deltas[name]
    all_ = alignments(d)
    one, many = oracle.delta_filter(d, 1), oracle.delta_filter(d, 2)
    a1, am = alignments(one), alignments(many)
    # -1 is the intersection, -m the union of the per-reference and per-query chains: subsets in input order
    it = iter(am); assert all(x in it for x in a1)
    it = iter(all_); assert all(x in it for x in am)
    if all_:
        assert a1 and am                                   # the best alignment of a sequence is on both of its chains... of some sequence
    # idempotent, header lines kept, no empty '>' blocks
    assert oracle.delta_filter(one, 1) == one and oracle.delta_filter(many, 2) == many
    assert one.split(b"\n")[:2] == d.split(b"\n")[:2]
    lines = one.split(b"\n")
    assert not any(lines[i].startswith(b">") and (lines[i + 1].startswith(b">") or lines[i + 1] == b"") for i in range(len(lines) - 1))
    assert hashlib.sha256(one).hexdigest() == GOLD[name]["filter1_sha256"] and hashlib.sha256(many).hexdigest() == GOLD[name]["filterm_sha256"]
    assert (len(all_), len(a1), len(am)) == tuple(GOLD[name]["n_alignments"])


def test_filter_removes_what_it_should(oracle, deltas):
    ref, qry, d = deltas["dup_in_query"]
    all_, a1, am = alignments(d), alignments(oracle.delta_filter(d, 1)), alignments(oracle.delta_filter(d, 2))
    assert len(a1) < len(am) <= len(all_)
    # after -1 no reference base is covered by two alignments for more than the allowed overlap
    spans = sorted((a[1][0], a[1][1]) for a in a1)
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        olap = e0 - s1 + 1
        assert olap <= 0 or (olap / (e0 - s0 + 1) <= 0.75 and olap / (e1 - s1 + 1) <= 0.75)


def test_filter_known_answer(oracle):
    # two alignments on the same reference range, the second one better: -1 keeps the second only;
    # a third elsewhere stays.  Scores: len * idy^2.
    d = (b"r.fa q.fa\nNUCMER\n>R Q 10000 10000\n"
         b"100 1099 100 1099 50 50 0\n0\n"          # idy 0.95
         b"100 1099 5000 5999 10 10 0\n0\n"         # idy 0.99: wins the reference chain
         b"3000 3999 7000 7999 0 0 0\n0\n")
    one = oracle.delta_filter(d, 1)
    assert [a[1][:4] for a in alignments(one)] == [(100, 1099, 5000, 5999), (3000, 3999, 7000, 7999)]
    assert len(alignments(oracle.delta_filter(d, 2))) == 3      # every one is on the query chain
    with pytest.raises(ValueError):
        oracle.delta_filter(b"r q\nNUCMER\n>R Q 10 10\n1 2 3\n0\n", 1)


def walk_maf(maf: bytes):
    blocks, cur = [], None
    lines = maf.split(b"\n")
    assert lines[0].startswith(b"##maf ")
    for l in lines[1:]:
        if l.startswith(b"a score="):
            cur = []
        elif l.startswith(b"s "):
            t = [x for x in re.split(rb"[ \t]", l) if x]      # lib/maf/reader.ml:21-27
            assert len(t) == 7 and t[4] in (b"+", b"-")
            cur.append((t[1].decode(), int(t[2]), int(t[3]), t[4].decode(), int(t[5]), t[6]))
        elif l == b"":
            if cur:
                blocks.append(cur)
            cur = None
        else:
            raise AssertionError(b"line the reference's readers reject: " + l[:60])   # lib/profiles/m_untranslate.ml:148-149
    return blocks


COMP = bytes.maketrans(b"ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", b"TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn")


@pytest.mark.parametrize("name", sorted(POST_CASES))
def test_maf_replays_the_sequences_and_golden(oracle, deltas, name):
    ref, qry, d = deltas[name]
    maf = oracle.delta2maf(d, ref, qry)
    blocks = walk_maf(maf)
    als = alignments(d)
    assert len(blocks) == len(als)
    R, Q = dict(H.parse_fasta(ref)), dict(H.parse_fasta(qry))
    for (hdr, a, dl), blk in zip(als, blocks):
        (rn, rs, rl, rd, rt, rtxt), (qn, qs, ql, qd, qt, qtxt) = blk
        rid, qid, rlen, qlen = hdr[1:].split()
        assert (rn, qn, rt, qt) == (rid.decode(), qid.decode(), int(rlen), int(qlen)) and rd == "+"
        assert len(rtxt) == len(qtxt) == rl + sum(1 for x in dl if x < 0)
        assert rtxt.replace(b"-", b"") == R[rn][rs:rs + rl] and rs == a[0] - 1 and rl == a[1] - a[0] + 1
        strand = Q[qn] if qd == "+" else Q[qn][::-1].translate(COMP)
        assert qtxt.replace(b"-", b"") == strand[qs:qs + ql] and ql == abs(a[3] - a[2]) + 1
        assert (qd == "-") == (a[2] > a[3])
        # column count of errors = the delta's error field (any non-identical-acgt column and every gap)
        errs = sum(1 for x, y in zip(rtxt.upper(), qtxt.upper()) if x != y or x not in b"ACGT")
        assert errs == a[4]
        assert not any(x == y == ord("-") for x, y in zip(rtxt, qtxt))
    assert hashlib.sha256(maf).hexdigest() == GOLD[name]["maf_sha256"]
    # delta2maf of the filtered delta (what mugsy_nucmer.ml:128-131 actually converts) holds the surviving blocks
    assert len(walk_maf(oracle.delta2maf(oracle.delta_filter(d, 1), ref, qry))) == len(alignments(oracle.delta_filter(d, 1)))


def test_filtered_delta_goes_through_the_reference_parser(oracle, deltas, tmp_path):
    if not os.path.exists(REF_TOOL):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    for name in ("dup_in_query", "shuffled_records", "100k_98_inv"):
        one = oracle.delta_filter(deltas[name][2], 1)
        p = tmp_path / (name + ".delta"); p.write_bytes(one)
        out = subprocess.run([REF_TOOL, str(p)], capture_output=True)
        assert out.returncode == 0, out.stderr
        # M_delta_stream -> M_delta_stream_writer re-encodes every alignment (metadata written as 1 2 3, m_delta_stream_writer.hh:71)
        got = [(a[1][:4], a[2]) for a in alignments(b"x y\nNUCMER\n" + out.stdout.split(b"\n", 2)[2])] if out.stdout.count(b"\n") > 2 else []
        assert got == [(a[1][:4], a[2]) for a in alignments(one)]


# ---- the reference's own consumer of these deltas at merge nodes (lib/m_translate, compiled as it is)

M_TRANSLATE = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "m_translate")


def run_m_translate(tmp_path, delta: bytes, ref: bytes, qry: bytes, split_ref_at=None):
    """Writes profile directories in the format of lib/profiles_lib/m_profile.cc:15-84 (one gap-free profile per
    record, or two per reference record when split_ref_at is given) and runs `m_translate left right list out`
    (lib/m_translate/m_translate_main.cc:22-45)."""
    (tmp_path / "in.delta").write_bytes(delta)
    for side, fa in (("left", ref), ("right", qry)):
        os.makedirs(tmp_path / side, exist_ok=True)
        with open(tmp_path / side / "profiles", "w") as f:
            for k, (name, seq) in enumerate(H.parse_fasta(fa)):
                cuts = [(1, len(seq))]
                if side == "left" and split_ref_at and len(seq) > split_ref_at:
                    cuts = [(1, split_ref_at), (split_ref_at + 1, len(seq))]
                for m, (s, e) in enumerate(cuts):
                    f.write(f"{side[0].upper()}{k} {m} {name} {s} {e} {e - s + 1} {len(seq)}\n0\n{seq[s - 1:e].decode()}\n")
    (tmp_path / "list").write_text(str(tmp_path / "in.delta") + "\n")
    out = subprocess.run([M_TRANSLATE, str(tmp_path / "left"), str(tmp_path / "right"), str(tmp_path / "list"), str(tmp_path / "out.delta")], capture_output=True)
    assert out.returncode == 0, out.stderr          # an assertion failure of m_translate.cc aborts with a non-zero status
    return (tmp_path / "out.delta").read_bytes()


@pytest.mark.parametrize("name", ["100k_98_inv", "multirecord", "shuffled_records", "big_indels", "10k_95"])
def test_deltas_drive_the_reference_m_translate(oracle, deltas, tmp_path, name):
    if not os.path.exists(M_TRANSLATE):
        pytest.skip("oracle/_ref/m_translate not built (no /root/reference here)")
    ref, qry, d = deltas[name]
    d = oracle.delta_filter(d, 1)                   # what reaches the merge nodes is the filtered delta
    # identity profiles: the translation must give back every alignment, coordinates and deltas unchanged
    got = alignments(run_m_translate(tmp_path, d, ref, qry))
    assert [(a[1][:4], a[2]) for a in got] == [(a[1][:4], a[2]) for a in alignments(d)]
    # the reference sequence cut into two profiles: alignments are split at the cut, nothing is lost
    cut = 4_000
    got2 = alignments(run_m_translate(tmp_path, d, ref, qry, split_ref_at=cut))
    assert len(got2) >= len(got)
    assert sum(a[1][1] - a[1][0] + 1 for a in got2) == sum(a[1][1] - a[1][0] + 1 for a in got)
