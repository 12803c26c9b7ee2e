"""Host logic mirroring lib/nucmer/mugsy_nucmer.ml and its callers (CPU)."""
import pytest

from paramugsy_b200 import mugsy_nucmer as M


def test_parse_argv_defaults_and_prefixing():
    o = M.parse_argv(["-ref_seq", "r.fa", "-query_seq", "q.fa", "-maf_out", "x.maf", "-delta_out", "x.delta"])
    # defaults of mugsy_nucmer.ml:47-58 and the out_dir prefix of :85-86
    assert (o.out_dir, o.tmp_dir, o.filter, o.colinear, o.debug, o.nucmer_opts, o.delta_pp) == ("/tmp", "/tmp", True, False, False, "", None)
    assert o.maf_out == "/tmp/x.maf" and o.delta_out == "/tmp/x.delta"
    o = M.parse_argv(["-out_dir", "/o", "-tmp_dir", "/o/r-q", "-ref_seq", "r", "-query_seq", "q", "-maf_out", "m", "-delta_out", "d",
                      "-nofilter", "-colinear", "-debug", "-delta_pp", "pp", "-nucmer_opts", "-l 15"])
    assert (o.filter, o.colinear, o.debug, o.delta_pp, o.nucmer_opts, o.delta_out) == (False, True, True, "pp", "-l 15", "/o/d")


def test_parse_argv_failures_match_reference_messages():
    with pytest.raises(M.Failure, match="Must specify -ref_seq and -query_seq"):      # mugsy_nucmer.ml:78-79
        M.parse_argv(["-maf_out", "m", "-delta_out", "d"])
    with pytest.raises(M.Failure, match="Must specify -maf_out and -delta_out"):      # mugsy_nucmer.ml:80-81
        M.parse_argv(["-ref_seq", "r", "-query_seq", "q"])
    with pytest.raises(M.Failure):
        M.parse_argv(["-bogus"])


def test_nucmer_opts_translation():
    o = M.nucmer_opts_to_pmn("-l 15 -c 40 -g 100 -D 7 -d 0.2 -b 150 -f --nosimplify --noextend")
    assert (o.minmatch, o.mincluster, o.maxgap, o.diagdiff, o.breaklen) == (15, 40, 100, 7, 150)
    assert abs(o.diagfactor - 0.2) < 1e-12 and (o.do_forward, o.do_reverse, o.do_simplify, o.do_extend) == (1, 0, 0, 0)
    with pytest.raises(M.Failure):
        M.nucmer_opts_to_pmn("--maxmatch")


def test_nucmer_task_naming_and_commands():
    s = [("/d/a.fa", "/d/b.fa"), ("/d/a.fa", "/e/c.fa")]
    assert M.basename(*s[0]) == "a.fa-b.fa"
    p = M.out_paths("/t", s)
    assert p == {"a.fa-b.fa-maf": "/t/a.fa-b.fa.maf", "a.fa-b.fa-delta": "/t/a.fa-b.fa.delta",
                 "a.fa-c.fa-maf": "/t/a.fa-c.fa.maf", "a.fa-c.fa-delta": "/t/a.fa-c.fa.delta"}
    assert M.make_commands(s[:1], "/t") == [
        "mugsy_nucmer -ref_seq /d/a.fa -query_seq /d/b.fa -out_dir /t -tmp_dir /t/a.fa-b.fa -maf_out a.fa-b.fa.maf -delta_out a.fa-b.fa.delta"]


def test_pm_job_pair_enumeration():
    g = ["g0", "g1", "g2", "g3"]
    assert M.searches(g) == [("g0", "g1"), ("g0", "g2"), ("g0", "g3"), ("g1", "g2"), ("g1", "g3"), ("g2", "g3")]
    assert M.cross(["a", "b"], ["x", "y"]) == [("a", "x"), ("a", "y"), ("b", "x"), ("b", "y")]
    assert len(M.searches([str(i) for i in range(8)])) == 28          # BASELINE.json configs[1]
    assert M.chunk(10, list(range(28))) == [list(range(10)), list(range(10, 20)), list(range(20, 28))]   # paramugsy.ml:34
    # configs[2]: 57 genomes split 7,7,7,7,7,7,7,8 by pm_job.mk_job -> 1596 = C(57,2) pairs in total
    def job(l):
        if len(l) <= 10: return len(M.searches(l))
        left, right = l[:len(l) // 2], l[len(l) // 2:]
        return job(left) + job(right) + len(M.cross(left, right))
    assert job(list(range(57))) == 57 * 56 // 2
