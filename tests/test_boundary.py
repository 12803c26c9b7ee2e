"""The drop-in boundary as the reference meets it:

  * the `nucmer` argv shim — the child process of lib/nucmer/mugsy_nucmer.ml:100,
    `nucmer <ref.fa> <qry.fa> -p <prefix> <opts>` -> `<prefix>.delta`, non-zero exit on any failure;
  * the OCaml stubs of integration/pmn_stubs.c, compiled against a stand-in for <caml/...> (tests/fake_caml) and driven
    from C in the order the patched mugsy_nucmer.ml calls them (integration/mugsy_nucmer.ml.patch);
  * the option table both share with the Python mirror (pmn_nucmer_parse_argv, pmn_opts_parse).

CPU tests cover everything up to the first device call (argument errors, missing files, "no CUDA device");
the -m gpu tests run the binaries end to end and compare bytes with the oracle."""
import os
import shutil
import subprocess

import pytest

from paramugsy_b200 import build, lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "paramugsy_b200", "_lib")


@pytest.fixture(scope="module")
def shim():
    build.build()
    return os.path.join(LIBDIR, "nucmer")


@pytest.fixture(scope="module")
def stub_driver(tmp_path_factory):
    build.build()
    exe = str(tmp_path_factory.mktemp("stub") / "driver")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-std=c11", "-D_GNU_SOURCE",
                           "-I" + os.path.join(ROOT, "tests", "fake_caml"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "integration", "pmn_stubs.c"), os.path.join(ROOT, "tests", "fake_caml", "driver.c"),
                           "-o", exe, "-L" + LIBDIR, "-lpmnucmer", "-Wl,-rpath," + LIBDIR])
    return exe


@pytest.fixture(scope="module")
def pair(tmp_path_factory):
    d = tmp_path_factory.mktemp("pair")
    g0 = synth.random_genome(60_000, 77)
    g1 = synth.invert(synth.mutate(g0, 0.04, 78), 1, 4_000, 79)
    ref, qry = synth.fasta("sp0.chr", g0), synth.fasta("sp1.chr", g1)
    (d / "sp0").write_bytes(ref); (d / "sp1").write_bytes(qry)
    return d, ref, qry


# ------------------------------------------------------------------ the shared option table (no GPU)

def test_option_table_reads_nucmer_spellings():
    o, prefix, ref, qry, dev, h, v = lib.nucmer_parse_argv(["r.fa", "q.fa", "-p", "tmp/nucmer", "-l", "15", "--mincluster=40", "-g", "100", "-D", "7",
                                                            "-d", "0.2", "--breaklen", "150", "-f", "--nosimplify", "--noextend", "--mumreference"])
    assert (prefix, ref, qry, dev, h, v) == ("tmp/nucmer", "r.fa", "q.fa", None, False, False)
    assert (o.minmatch, o.mincluster, o.maxgap, o.diagdiff, o.diagfactor, o.breaklen) == (15, 40, 100, 7, 0.2, 150)
    assert (o.do_forward, o.do_reverse, o.do_simplify, o.do_extend, o.do_optimize) == (1, 0, 0, 0, 1)
    d = lib.default_opts()
    e = lib.opts_from_nucmer_string("")            # lib/base/nucmer_task.ml:53 never sets -nucmer_opts
    assert all(getattr(d, k) == getattr(e, k) for k, _ in lib.Opts._fields_)
    assert lib.opts_from_nucmer_string("  -b   300  '-c' \"70\" ").breaklen == 300


@pytest.mark.parametrize("bad, msg", [("--maxmatch", "not implemented"), ("--mum", "not implemented"), ("--nooptimize", "not implemented"),
                                      ("--frobnicate", "unknown option"), ("-l", "needs a value"), ("-l x", "needs an integer"),
                                      ("-f -r", "mutually exclusive"), ("-p elsewhere", "alignment options only"), ("extra.fa", "alignment options only"),
                                      ("'-l 15", "unbalanced quote")])
def test_option_table_rejects(bad, msg):
    with pytest.raises(lib.PmnError) as e:
        lib.opts_from_nucmer_string(bad)
    assert e.value.code == -1 and msg in str(e.value)


# ------------------------------------------------------------------ the `nucmer` shim

def test_shim_argument_errors_exit_1_before_any_device_call(shim, pair, tmp_path):
    d, _, _ = pair
    for argv, msg in ((["--bogus", str(d / "sp0"), str(d / "sp1")], "unknown option --bogus"),
                      ([str(d / "sp0"), str(d / "sp1"), "--maxmatch"], "not implemented"),
                      ([str(d / "sp0")], "USAGE"),
                      ([str(d / "sp0"), str(tmp_path / "missing.fa")], "cannot open")):
        p = subprocess.run([shim] + argv, capture_output=True, text=True, cwd=tmp_path)
        assert p.returncode == 1 and msg in p.stderr, (argv, p.returncode, p.stderr)
        assert not os.path.exists(tmp_path / "out.delta")
    p = subprocess.run([shim, "-h"], capture_output=True, text=True)
    assert p.returncode == 0 and "USAGE: nucmer" in p.stdout


@pytest.mark.gpu
def test_shim_writes_prefix_delta_like_the_child_process(shim, pair, tmp_path, oracle):
    """mugsy_nucmer.ml:96-100: `nucmer R Q -p <tmp>/nucmer` must leave <tmp>/nucmer.delta; line 1 carries absolute paths."""
    d, ref, qry = pair
    tmp = tmp_path / "tmpdir"; tmp.mkdir()
    p = subprocess.run([shim, "sp0", "sp1", "-p", str(tmp / "nucmer")], capture_output=True, text=True, cwd=d)
    assert p.returncode == 0, p.stderr
    got = (tmp / "nucmer.delta").read_bytes()
    r0, r1 = os.path.realpath(d / "sp0"), os.path.realpath(d / "sp1")
    assert got == oracle.nucmer(ref, qry, r0, r1)
    assert os.listdir(tmp) == ["nucmer.delta"]                       # written atomically: no temporary file left
    # options reach the library: the same call with -l 30 -b 100 equals the oracle run with those options
    p = subprocess.run([shim, str(d / "sp0"), str(d / "sp1"), "--prefix=" + str(tmp / "o2"), "-l", "30", "-b", "100"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert (tmp / "o2.delta").read_bytes() == oracle.nucmer(ref, qry, r0, r1, minmatch=30, breaklen=100)
    # the two post-step front ends pick the FASTA files up from line 1, from any working directory (mugsy_profiles_task.ml:60)
    f = subprocess.run([os.path.join(LIBDIR, "delta-filter"), "-1", str(tmp / "nucmer.delta")], capture_output=True, cwd=tmp_path)
    assert f.returncode == 0 and f.stdout == oracle.delta_filter(got, 1)
    (tmp / "filt.delta").write_bytes(f.stdout)
    m = subprocess.run([os.path.join(LIBDIR, "delta2maf"), str(tmp / "filt.delta")], capture_output=True, cwd=tmp_path)
    assert m.returncode == 0 and m.stdout == oracle.delta2maf(f.stdout, ref, qry)


# ------------------------------------------------------------------ the OCaml stubs

def test_stub_raises_failure_with_the_lock_held(stub_driver, pair, tmp_path):
    d, _, _ = pair
    p = subprocess.run([stub_driver, str(d / "sp0"), str(d / "sp1"), "--maxmatch", str(tmp_path), str(tmp_path / "o.delta"), str(tmp_path / "o.maf")],
                       capture_output=True, text=True)
    assert p.returncode == 2 and "Failure: nucmer: option --maxmatch is not implemented" in p.stdout, (p.returncode, p.stdout, p.stderr)
    if lib.lib().pmn_device_count() == 0:
        p = subprocess.run([stub_driver, str(d / "sp0"), str(d / "sp1"), "", str(tmp_path), str(tmp_path / "o.delta"), str(tmp_path / "o.maf")],
                           capture_output=True, text=True)
        assert p.returncode == 2 and "no CUDA device" in p.stdout          # no CPU fallback behind the stub either


def test_patch_applies_to_the_reference_worker(tmp_path):
    ref = "/root/reference/lib/nucmer"
    if not os.path.isdir(ref) or shutil.which("patch") is None:
        pytest.skip("the reference tree is only present in the build container")
    dst = tmp_path / "lib" / "nucmer"; dst.mkdir(parents=True)
    for f in ("mugsy_nucmer.ml", "Makefile"):
        shutil.copy(os.path.join(ref, f), dst / f)
    subprocess.check_call(["patch", "-p1", "-s", "-i", os.path.join(ROOT, "integration", "mugsy_nucmer.ml.patch")], cwd=tmp_path)
    ml = (dst / "mugsy_nucmer.ml").read_text()
    assert 'Shell.sh' not in ml.split("let nucmer options")[1].split("match options.delta_pp")[0]      # nucmer and delta-filter no longer fork
    assert "pmn_align_pair ref_file query_file options.nucmer_opts delta_file" in ml and "pmn_stubs.c" in (dst / "Makefile").read_text()


@pytest.mark.gpu
@pytest.mark.parametrize("flags, mode", [([], 1), (["-colinear"], 2), (["-nofilter"], 0)])
def test_stub_runs_the_worker_like_the_patched_mugsy_nucmer(stub_driver, pair, tmp_path, oracle, flags, mode):
    d, ref, qry = pair
    tmp = tmp_path / "t"; tmp.mkdir()
    a, b = str(d / "sp0"), str(d / "sp1")
    p = subprocess.run([stub_driver, a, b, "-b 150", str(tmp), str(tmp_path / "o.delta"), str(tmp_path / "o.maf")] + flags, capture_output=True, text=True)
    assert p.returncode == 0, (p.stdout, p.stderr)
    want = oracle.nucmer(ref, qry, a, b, breaklen=150)
    assert (tmp / "nucmer.delta").read_bytes() == want
    final = oracle.delta_filter(want, mode) if mode else want
    assert (tmp_path / "o.delta").read_bytes() == final
    assert (tmp_path / "o.maf").read_bytes() == oracle.delta2maf(final, ref, qry)
