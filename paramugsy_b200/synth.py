"""Deterministic synthetic genomes — ctypes binding of csrc/pmn_synth.c."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_lib", "libpmn_synth.so")
        if not os.path.exists(path):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", path,
                                   os.path.join(_HERE, "csrc", "pmn_synth.c")])
        L = C.CDLL(path)
        L.pmn_synth_random.argtypes = [C.c_char_p, C.c_int64, C.c_uint64]
        L.pmn_synth_mutate.argtypes = [C.c_char_p, C.c_int64, C.c_double, C.c_uint64, C.c_char_p]
        L.pmn_synth_mutate.restype = C.c_int64
        L.pmn_synth_invert.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int64, C.c_uint64]
        L.pmn_synth_invert.restype = C.c_int
        _LIB = L
    return _LIB


def random_genome(n: int, seed: int) -> bytes:
    buf = C.create_string_buffer(n)
    _lib().pmn_synth_random(buf, n, seed)
    return buf.raw


def mutate(g: bytes, d: float, seed: int) -> bytes:
    out = C.create_string_buffer(2 * len(g) + 16)
    m = _lib().pmn_synth_mutate(g, len(g), d, seed, out)
    return out.raw[:m]


def invert(g: bytes, k: int, length: int, seed: int) -> bytes:
    buf = C.create_string_buffer(g, len(g))
    _lib().pmn_synth_invert(buf, len(g), k, length, seed)
    return buf.raw


def fasta(name: str, seq: bytes, width: int = 60) -> bytes:
    """One FASTA record; `name` should look like species.accession (m_rewrite_fasta.ml:5-59)."""
    lines = [b">" + name.encode()]
    lines += [seq[i:i + width] for i in range(0, len(seq), width)]
    return b"\n".join(lines) + b"\n"


# --- the benchmark configurations of BASELINE.json / SURVEY.md §8d -------------------------

def config_c1(n=1_000_000):
    g0 = random_genome(n, 1001)
    return [("g0.1", g0), ("g1.1", mutate(g0, 0.01, 1002))]


def config_c2(n=5_000_000, count=8, inv_len=50_000):
    anc = random_genome(n, 2000)
    return [(f"g{i}.1", invert(mutate(anc, 0.02, 2001 + i), 2, inv_len, 2101 + i)) for i in range(count)]


def config_c3(n=2_000_000, count=57):
    anc = random_genome(n, 3000)
    return [(f"s{i}.1", mutate(anc, 0.03, 3001 + i)) for i in range(count)]


def config_c4(n=100_000_000, inv_len=1_000_000):
    g0 = random_genome(n, 4000)
    return [("c0.1", g0), ("c1.1", invert(mutate(g0, 0.01, 4001), 8, inv_len, 4101))]


def config_c5(n=5_000_000, ds=(0.01, 0.02, 0.03, 0.05, 0.08, 0.10, 0.12, 0.15)):
    anc = random_genome(n, 5000)
    return ("anc.1", anc), [(f"q{int(round(d * 100)):02d}.1", mutate(anc, d, 5000 + int(1000 * d))) for d in ds]


def searches(genomes):
    """Upper-triangle ordered pairs, earlier genome = reference (lib/base/pm_job.ml:43-51)."""
    return [(genomes[i], genomes[j]) for i in range(len(genomes)) for j in range(i + 1, len(genomes))]
