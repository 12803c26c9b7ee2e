"""GPU parity of the two post-steps (-m gpu): pmn_delta_filter / pmn_delta2maf through the C ABI, the argv
front ends and the mugsy_nucmer mirror, against the CPU oracle (oracle/pmn_post_oracle.c) and the golden digests."""
import hashlib
import json
import os
import subprocess

import pytest

from post_cases import POST_CASES
from paramugsy_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_post.json")))
LIBDIR = os.path.join(os.path.dirname(__file__), "..", "paramugsy_b200", "_lib")


@pytest.fixture(scope="module")
def ctx():
    from paramugsy_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", sorted(POST_CASES))
def test_filter_and_maf_equal_oracle_and_golden(ctx, oracle, name):
    ref, qry, kw = POST_CASES[name]()
    rs, qs = ctx.sequence(ref), ctx.sequence(qry)
    ix = rs.index()
    res = ix.align(qs, ref_path="ref.fa", qry_path="qry.fa", **kw)
    d = res.delta
    res.close(); ix.close()
    assert hashlib.sha256(d).hexdigest() == GOLD[name]["delta_sha256"]
    one, many = ctx.delta_filter(d, 1), ctx.delta_filter(d, 2)
    assert one == oracle.delta_filter(d, 1) and many == oracle.delta_filter(d, 2)
    assert hashlib.sha256(one).hexdigest() == GOLD[name]["filter1_sha256"] and hashlib.sha256(many).hexdigest() == GOLD[name]["filterm_sha256"]
    maf = ctx.delta2maf(d, rs, qs)
    assert maf == oracle.delta2maf(d, ref, qry)
    assert hashlib.sha256(maf).hexdigest() == GOLD[name]["maf_sha256"]
    assert ctx.delta2maf(one, rs, qs) == oracle.delta2maf(one, ref, qry)        # what mugsy_nucmer.ml:128-131 converts
    assert ctx.delta_filter(d, 1, 10.0) == oracle.delta_filter(d, 1, 10.0)     # delta-filter -o
    qs.close(); rs.close()


def test_filter_many_alignments_per_sequence(ctx, oracle):
    """A synthetic .delta with hundreds of overlapping alignments on few sequences: the chain DP proper."""
    import random
    rnd = random.Random(5)
    lines = ["r.fa q.fa", "NUCMER"]
    for rname, qname in (("R1", "Q1"), ("R1", "Q2"), ("R2", "Q1")):
        lines.append(f">{rname} {qname} 2000000 2000000")
        for _ in range(300):
            s = rnd.randrange(1, 1_900_000); ln = rnd.randrange(200, 60_000); e = rnd.randrange(0, ln // 8)
            qs_ = rnd.randrange(1, 1_900_000); rev = rnd.random() < 0.3
            q = (qs_ + ln - 1, qs_) if rev else (qs_, qs_ + ln - 1)
            lines.append(f"{s} {s + ln - 1} {q[0]} {q[1]} {e} {e} 0")
            pos = 0
            for _ in range(rnd.randrange(0, 4)):            # a few indels that keep both lengths equal
                lines.append(str(rnd.randrange(2, 50))); lines.append(str(-rnd.randrange(2, 50)))
            lines.append("0")
    d = ("\n".join(lines) + "\n").encode()
    for mode in (1, 2):
        assert ctx.delta_filter(d, mode) == oracle.delta_filter(d, mode)
    assert len(ctx.delta_filter(d, 1)) < len(ctx.delta_filter(d, 2)) < len(d)


def test_maf_full_size_pair_properties(ctx):
    """BASELINE.json's pair size (5 Mbp), where the oracle's MAF would take a while in the test: both rows of
    every block have the same length, stripped of gaps they are the aligned stretches of the two genomes."""
    gs = synth.config_c2(count=2)
    ref, qry = synth.fasta(*gs[0]), synth.fasta(*gs[1])
    rs, qs = ctx.sequence(ref), ctx.sequence(qry)
    ix = rs.index(); res = ix.align(qs, ref_path="r", qry_path="q"); d = res.delta; res.close(); ix.close()
    one = ctx.delta_filter(d, 1)
    maf = ctx.delta2maf(one, rs, qs)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    R, Q = gs[0][1], gs[1][1]
    blocks = maf.split(b"\n\n")
    assert blocks[0].startswith(b"##maf version=1\na score=0\n")
    n = 0; covered = 0
    for b in blocks:
        rows = [l.split(b" ") for l in b.split(b"\n") if l.startswith(b"s ")]
        if not rows:
            continue
        (_, rn, rstart, rsize, rd, rtot, rtxt), (_, qn, qstart, qsize, qd, qtot, qtxt) = rows
        assert len(rtxt) == len(qtxt) and rd == b"+" and int(rtot) == len(R) and int(qtot) == len(Q)
        assert rtxt.replace(b"-", b"") == R[int(rstart):int(rstart) + int(rsize)]
        strand = Q if qd == b"+" else Q[::-1].translate(comp)
        assert qtxt.replace(b"-", b"") == strand[int(qstart):int(qstart) + int(qsize)]
        n += 1; covered += int(rsize)
    assert n == one.count(b"\n0\n") and covered > 0.97 * len(R)
    qs.close(); rs.close()


def test_post_error_paths(ctx):
    from paramugsy_b200 import lib
    rs = ctx.sequence(b">a\nACGTACGTACGTACGTACGTAAAACCCCGGGG\n")
    with pytest.raises(lib.PmnError):
        ctx.delta_filter(b"r q\nNUCMER\n>a a 32 32\n1 2 3\n0\n", 1)           # six fields missing
    with pytest.raises(lib.PmnError):
        ctx.delta_filter(b"r q\nNUCMER\n>a a 32 32\n1 10 1 10 0 0 0\n5\n", 1)    # no terminating 0
    with pytest.raises(lib.PmnError):
        ctx.delta2maf(b"r q\nNUCMER\n>zzz a 32 32\n1 10 1 10 0 0 0\n0\n", rs, rs)   # sequence not in the FASTA
    with pytest.raises(lib.PmnError):
        ctx.delta2maf(b"r q\nNUCMER\n>a a 32 32\n1 40 1 40 0 0 0\n0\n", rs, rs)    # beyond the sequence
    with pytest.raises(lib.PmnError):
        ctx.delta2maf(b"r q\nNUCMER\n>a a 32 32\n1 10 1 12 0 0 0\n0\n", rs, rs)    # deltas do not account for the lengths
    assert ctx.delta2maf(b"r q\nNUCMER\n", rs, rs) == b"##maf version=1\n"
    assert ctx.delta_filter(b"r q\nNUCMER\n", 1) == b"r q\nNUCMER\n"
    assert ctx.delta2maf(b"r q\nNUCMER\n>a a 32 32\n1 10 1 10 0 0 0\n0\n", rs, rs) == b"##maf version=1\na score=0\ns a 0 10 + 32 ACGTACGTAC\ns a 0 10 + 32 ACGTACGTAC\n\n"
    rs.close()


def test_front_ends_and_mirror(oracle, tmp_path):
    """The argv front ends (delta-filter -1 f > out, delta2maf f > out) and the mugsy_nucmer mirror with its
    default filter (lib/nucmer/mugsy_nucmer.ml:54,96-131): nucmer.delta, nucmer.filt.delta, delta_out, maf_out."""
    from paramugsy_b200 import mugsy_nucmer as M
    ref, qry, _ = POST_CASES["shuffled_records"]()
    rp, qp = tmp_path / "ref.fa", tmp_path / "qry.fa"
    rp.write_bytes(ref); qp.write_bytes(qry)
    d = oracle.nucmer(ref, qry, str(rp), str(qp), fast_chain=1)
    dp = tmp_path / "in.delta"; dp.write_bytes(d)
    for mode, flag in ((1, "-1"), (2, "-m")):
        out = subprocess.run([os.path.join(LIBDIR, "delta-filter"), flag, str(dp)], capture_output=True)
        assert out.returncode == 0, out.stderr
        assert out.stdout == oracle.delta_filter(d, mode)
    out = subprocess.run([os.path.join(LIBDIR, "delta2maf"), str(dp)], capture_output=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout == oracle.delta2maf(d, ref, qry)
    assert subprocess.run([os.path.join(LIBDIR, "delta-filter"), str(dp)], capture_output=True).returncode == 1
    assert subprocess.run([os.path.join(LIBDIR, "delta2maf"), str(tmp_path / "nope")], capture_output=True).returncode == 1
    # the mirror
    outd, tmpd = tmp_path / "out", tmp_path / "tmp"
    o = M.parse_argv(["-ref_seq", str(rp), "-query_seq", str(qp), "-maf_out", "p.maf", "-delta_out", "p.delta", "-out_dir", str(outd), "-tmp_dir", str(tmpd), "-debug"])
    os.makedirs(outd); os.makedirs(tmpd)
    M.run_search(o)
    want = oracle.delta_filter(oracle.nucmer(ref, qry, str(rp), str(qp), fast_chain=1), 1)
    assert (tmpd / "nucmer.filt.delta").read_bytes() == want == (outd / "p.delta").read_bytes()
    assert (outd / "p.maf").read_bytes() == oracle.delta2maf(want, ref, qry)
    o2 = M.parse_argv(["-ref_seq", str(rp), "-query_seq", str(qp), "-maf_out", "c.maf", "-delta_out", "c.delta", "-out_dir", str(outd), "-tmp_dir", str(tmpd), "-colinear"])
    M.run_search(o2)
    assert (outd / "c.delta").read_bytes() == oracle.delta_filter(oracle.nucmer(ref, qry, str(rp), str(qp), fast_chain=1), 2)
    # -delta_pp (lib/nucmer/mugsy_nucmer.ml:107-114): the filtered delta is piped through the post-processor into nucmer.pp.delta,
    # which is what delta_out and the MAF come from; a program that does not exist is a Failure
    o3 = M.parse_argv(["-ref_seq", str(rp), "-query_seq", str(qp), "-maf_out", "pp.maf", "-delta_out", "pp.delta", "-out_dir", str(outd), "-tmp_dir", str(tmpd),
                       "-delta_pp", "head -c 100000000"])
    M.run_search(o3)
    assert (tmpd / "nucmer.pp.delta").read_bytes() == want == (outd / "pp.delta").read_bytes()
    assert (outd / "pp.maf").read_bytes() == oracle.delta2maf(want, ref, qry)
    o4 = M.parse_argv(["-ref_seq", str(rp), "-query_seq", str(qp), "-maf_out", "x.maf", "-delta_out", "x.delta", "-out_dir", str(outd), "-tmp_dir", str(tmpd),
                       "-delta_pp", "no-such-delta-post-processor"])
    with pytest.raises(M.Failure):
        M.run_search(o4)
    # main (lib/nucmer/mugsy_nucmer.ml:134-140): makes the directories, runs the search, removes tmp_dir
    out2, tmp2 = tmp_path / "out2", tmp_path / "tmp2"
    M.main(["-ref_seq", str(rp), "-query_seq", str(qp), "-maf_out", "m.maf", "-delta_out", "m.delta", "-out_dir", str(out2), "-tmp_dir", str(tmp2)])
    assert (out2 / "m.delta").read_bytes() == want and (out2 / "m.maf").read_bytes() == oracle.delta2maf(want, ref, qry)
    assert not tmp2.exists()


def test_post_steps_inside_the_batch_call(oracle):
    """pmn_opts.post: nucmer, delta-filter and delta2maf of every pair in one scheduler call (what one mugsy_nucmer
    process does, lib/nucmer/mugsy_nucmer.ml:127-131) equal the three text-level steps and the oracle."""
    from paramugsy_b200 import lib
    gs = synth.config_c2(n=120_000, count=4, inv_len=3_000)
    dup = gs[1][1] + synth.random_genome(300, 77) + gs[1][1][40_000:60_000]         # a second copy: something for the filter to drop
    gs = gs + [("dup.1", dup)]
    fastas = [synth.fasta(n, s) for n, s in gs]; names = [n for n, _ in gs]
    pairs = [(i, j) for i in range(len(gs)) for j in range(i + 1, len(gs))]
    for mode in (1, 2):
        with lib.Scheduler(0, 3) as s:
            res = s.align_fasta(fastas, pairs, names=names, post=mode)
            dropped = 0
            for (i, j), r in zip(pairs, res):
                d = oracle.nucmer(fastas[i], fastas[j], names[i], names[j], fast_chain=1)
                f = oracle.delta_filter(d, mode)
                assert r.delta == d and r.filtered == f and r.maf == oracle.delta2maf(f, fastas[i], fastas[j])
                assert r.stats["wall_ms_post"] > 0
                dropped += d.count(b"\n0\n") - f.count(b"\n0\n")
            assert (dropped > 0) == (mode == 1)
    with lib.Scheduler(0, 2) as s:
        r = s.align_fasta(fastas[:2], [(0, 1)], names=names[:2])[0]
        assert r.filtered == b"" and r.maf == b""
        with pytest.raises(lib.PmnError):
            s.align_fasta(fastas[:2], [(0, 1)], post=3)


def test_run_nucmers_leaves_the_files_of_the_reference_workers(oracle, tmp_path):
    """The mirror of run_nucmers (lib/base/job_processor.ml:128-154) over pmn_worker_batch: per pair <bname>.delta is the
    filtered delta and <bname>.maf its MAF, named as lib/base/nucmer_task.ml:10-23 names them."""
    from paramugsy_b200 import mugsy_nucmer as M
    gs = synth.config_c2(n=90_000, count=3, inv_len=2_000)
    paths = []
    for name, seq in gs:
        p = tmp_path / name; p.write_bytes(synth.fasta(name, seq)); paths.append(str(p))
    searches = M.searches(paths)
    out = M.run_nucmers(searches, str(tmp_path / "nuc"))
    assert len(out) == 2 * len(searches)
    for a, b in searches:
        fa, fb = open(a, "rb").read(), open(b, "rb").read()
        f = oracle.delta_filter(oracle.nucmer(fa, fb, a, b, fast_chain=1), 1)
        bn = M.basename(a, b)
        assert open(out[bn + "-delta"], "rb").read() == f
        assert open(out[bn + "-maf"], "rb").read() == oracle.delta2maf(f, fa, fb)
    out2 = M.run_nucmers(searches[:1], str(tmp_path / "nuc2"), filter=False)
    a, b = searches[0]
    assert open(out2[M.basename(a, b) + "-delta"], "rb").read() == oracle.nucmer(open(a, "rb").read(), open(b, "rb").read(), a, b, fast_chain=1)
