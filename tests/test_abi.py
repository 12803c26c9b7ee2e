"""CPU checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/pmnucmer.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pmnucmer.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmn_[a-z_0-9]+)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    from paramugsy_b200 import build, lib
    so = build.build()
    assert os.path.exists(so)
    L = C.CDLL(so)
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/pmnucmer.h but not exported"
    assert sorted(lib.SYMBOLS) == declared, "paramugsy_b200/lib.py must bind exactly the declared symbols"


def test_kernels_are_sm_100a_only():
    from paramugsy_b200 import lib
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "--list-elf", lib.lib_path()]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_seeding_kernel_uses_tma_bulk_copy():
    from paramugsy_b200 import lib
    sass = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-sass", lib.lib_path()]).decode()
    blocks = [b for b in sass.split("Function : ")[1:] if b.startswith("_Z6k_seed")]
    assert len(blocks) == 1
    assert "UBLKCP" in blocks[0], "k_seed must stage its query tile with cp.async.bulk (UBLKCP in SASS)"
    assert "SYNCS.ARRIVE.TRANS64" in blocks[0], "expect_tx on the tile's mbarrier"


def test_no_gpu_means_error_not_fallback():
    from paramugsy_b200 import lib
    L = lib.lib()
    if L.pmn_device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(lib.PmnError) as e:
        lib.Context(0)
    assert e.value.code == -5 and "no CPU path" in str(e.value)


def test_default_opts_are_nucmer_defaults():
    from paramugsy_b200 import lib
    o = lib.default_opts()
    assert (o.minmatch, o.mincluster, o.maxgap, o.diagdiff, o.breaklen) == (20, 65, 90, 5, 200)
    assert abs(o.diagfactor - 0.12) < 1e-12
    assert (o.do_forward, o.do_reverse, o.do_extend, o.do_optimize, o.do_simplify) == (1, 1, 1, 1, 1)
    with pytest.raises(TypeError):
        lib.default_opts(nonsense=1)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under paramugsy_b200/ may import, link or run it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "paramugsy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                txt = re.sub(r"//[^\n]*|/\*.*?\*/", "", txt, flags=re.S)      # comments may cite the oracle
                if re.search(r"#\s*include[^\n]*oracle|libpmn_oracle|from\s+oracle|import\s+oracle|\bpmo_\w+\s*\(|oracle/", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_library_asks_for_a_hardware_queue_per_stream():
    """A scheduler's workers launch on 2 W streams; with the driver's default of 8 connections they alias and serialise
    (DESIGN.md §6).  The library sets the variable when it is loaded (a host that set it keeps its value), the binding when
    it is imported."""
    code = ("import os, ctypes; os.environ.pop('CUDA_DEVICE_MAX_CONNECTIONS', None); "
            "from paramugsy_b200 import build; ctypes.CDLL(build.build()); "
            "libc = ctypes.CDLL(None); libc.getenv.restype = ctypes.c_char_p; print(libc.getenv(b'CUDA_DEVICE_MAX_CONNECTIONS'))")
    out = subprocess.check_output([os.sys.executable, "-c", code], cwd=ROOT).decode()
    assert "b'32'" in out, out
    code = ("import os; os.environ['CUDA_DEVICE_MAX_CONNECTIONS'] = '4'; import paramugsy_b200.lib as l; l.lib(); "
            "print(os.environ['CUDA_DEVICE_MAX_CONNECTIONS'])")
    assert subprocess.check_output([os.sys.executable, "-c", code], cwd=ROOT).decode().strip() == "4"


def test_small_numbers_are_formatted_like_printf():
    """The .delta writers' integer formatter (pmn_host.h: branch-free below 10000, two digits per step above) against printf."""
    src = open(os.path.join(ROOT, "paramugsy_b200", "csrc", "pmn_host.h")).read()
    a = src.index("static const char PMN_DIGITS2[201]"); b = src.index("void pmn_apply_device_sched(int workers);")
    prog = ("#include <cstdio>\n#include <cstring>\n#include <cstdlib>\n#include <climits>\n#include <cstdint>\n" + src[a:b] + """
int main() {
    long long edge[] = {0, 1, -1, 9, 10, 99, 100, 999, 1000, 9999, -9999, 10000, -10000, 10001, 99999, 100000, 2147483647LL, -2147483648LL, LLONG_MAX, LLONG_MIN};
    char x[64], y[64]; int bad = 0;
    for (long long v : edge) { memset(x, '#', 64); *pmn_fmt_int(x, v) = 0; sprintf(y, "%lld", v); bad += strcmp(x, y) != 0; }
    for (long long v = -20000; v <= 20000; v++) { *pmn_fmt_int(x, v) = 0; sprintf(y, "%lld", v); bad += strcmp(x, y) != 0; }
    srand(7);
    for (int i = 0; i < 200000; i++) { long long v = ((long long)rand() << (rand() % 33)) - (1LL << (rand() % 40)); *pmn_fmt_int(x, v) = 0; sprintf(y, "%lld", v); bad += strcmp(x, y) != 0; }
    printf("%d\\n", bad); return bad != 0;
}
""")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(prog)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", os.path.join(d, "t"), os.path.join(d, "t.cpp")])
        assert subprocess.check_output([os.path.join(d, "t")]).decode().strip() == "0"
