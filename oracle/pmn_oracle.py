"""ctypes binding of the CPU oracle (oracle/libpmn_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under paramugsy_b200/ does.  PARITY UNPINNED vs MUMmer 3.20
(see pmn_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Opts(C.Structure):
    _fields_ = [("minmatch", C.c_int32), ("mincluster", C.c_int32), ("maxgap", C.c_int32),
                ("diagdiff", C.c_int32), ("diagfactor", C.c_double), ("breaklen", C.c_int32),
                ("do_forward", C.c_int32), ("do_reverse", C.c_int32), ("do_extend", C.c_int32),
                ("do_optimize", C.c_int32), ("do_simplify", C.c_int32), ("fast_chain", C.c_int32)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libpmn_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.pmo_run_create.restype = C.c_void_p
        L.pmo_run_create.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(Opts)]
        L.pmo_run_free.argtypes = [C.c_void_p]
        L.pmo_last_error.restype = C.c_char_p
        for name in ("pmo_stage_index", "pmo_stage_seed", "pmo_stage_cluster", "pmo_stage_extend"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.pmo_stage_delta.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        for name in ("pmo_ref_len", "pmo_n_anchors", "pmo_n_clusters", "pmo_n_cluster_matches",
                     "pmo_n_alignments", "pmo_dp_cells"):
            f = getattr(L, name); f.restype = C.c_int64; f.argtypes = [C.c_void_p]
        for name in ("pmo_ref_codes", "pmo_sa", "pmo_lcp", "pmo_anchors", "pmo_cluster_matches",
                     "pmo_cluster_off", "pmo_cluster_tag", "pmo_alignments", "pmo_delta_off", "pmo_deltas"):
            f = getattr(L, name); f.restype = C.c_void_p; f.argtypes = [C.c_void_p]
        L.pmo_delta_text.restype = C.c_void_p
        L.pmo_delta_text.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.pmo_default_opts.argtypes = [C.POINTER(Opts)]
        L.pmo_delta_filter.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_double, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.pmo_delta2maf.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.pmo_free.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def default_opts(**kw):
    o = Opts()
    lib().pmo_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class Run:
    """One (reference FASTA, query FASTA) pair pushed through the oracle stage by stage."""

    def __init__(self, ref_fasta: bytes, qry_fasta: bytes, opts=None, **kw):
        self.L = lib()
        self.opts = opts if opts is not None else default_opts(**kw)
        self.h = self.L.pmo_run_create(ref_fasta, len(ref_fasta), qry_fasta, len(qry_fasta), C.byref(self.opts))
        if not self.h:
            raise ValueError(self.L.pmo_last_error().decode())

    def close(self):
        if self.h:
            self.L.pmo_run_free(self.h)
            self.h = None

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.pmo_last_error().decode())

    def index(self):
        self._chk(self.L.pmo_stage_index(self.h))
        n = self.L.pmo_ref_len(self.h)
        return _arr(self.L.pmo_sa(self.h), n, np.int32), _arr(self.L.pmo_lcp(self.h), n, np.int32)

    def ref_codes(self):
        return _arr(self.L.pmo_ref_codes(self.h), self.L.pmo_ref_len(self.h), np.uint8)

    def anchors(self):
        """(n,4) int32: ref pos (1-based, concatenated), query pos (1-based, strand coords), len, tag."""
        self._chk(self.L.pmo_stage_seed(self.h))
        n = self.L.pmo_n_anchors(self.h)
        return _arr(self.L.pmo_anchors(self.h), n * 4, np.int32).reshape(n, 4)

    def clusters(self):
        """(matches (m,3) int32, off (k+1,) int32, tag (k,) int32) in mgaps output order."""
        self._chk(self.L.pmo_stage_cluster(self.h))
        k = self.L.pmo_n_clusters(self.h); m = self.L.pmo_n_cluster_matches(self.h)
        return (_arr(self.L.pmo_cluster_matches(self.h), m * 3, np.int32).reshape(m, 3),
                _arr(self.L.pmo_cluster_off(self.h), k + 1, np.int32),
                _arr(self.L.pmo_cluster_tag(self.h), k, np.int32))

    def alignments(self):
        """(rows (a,10) int64, delta_off (a+1,) int64, deltas int64)."""
        self._chk(self.L.pmo_stage_extend(self.h))
        a = self.L.pmo_n_alignments(self.h)
        off = _arr(self.L.pmo_delta_off(self.h), a + 1, np.int64)
        return (_arr(self.L.pmo_alignments(self.h), a * 10, np.int64).reshape(a, 10), off,
                _arr(self.L.pmo_deltas(self.h), int(off[-1]) if a else 0, np.int64))

    def delta(self, ref_path="ref.fa", qry_path="qry.fa") -> bytes:
        self._chk(self.L.pmo_stage_delta(self.h, ref_path.encode(), qry_path.encode()))
        n = C.c_size_t()
        p = self.L.pmo_delta_text(self.h, C.byref(n))
        return C.string_at(p, n.value)

    def dp_cells(self):
        return self.L.pmo_dp_cells(self.h)


def nucmer(ref_fasta: bytes, qry_fasta: bytes, ref_path="ref.fa", qry_path="qry.fa", **kw) -> bytes:
    r = Run(ref_fasta, qry_fasta, **kw)
    try:
        return r.delta(ref_path, qry_path)
    finally:
        r.close()


def _text_out(rc, p, n, what):
    if rc != 0:
        raise ValueError(f"{what}: malformed input")
    try:
        return C.string_at(p.value, n.value) if n.value else b""
    finally:
        lib().pmo_free(p)


def delta_filter(delta: bytes, mode: int = 1, maxolap: float = 75.0) -> bytes:
    """`delta-filter -1` (mode 1) / `-m` (mode 2) on .delta text (lib/nucmer/mugsy_nucmer.ml:102-105)."""
    p, n = C.c_void_p(), C.c_size_t()
    return _text_out(lib().pmo_delta_filter(delta, len(delta), mode, maxolap, C.byref(p), C.byref(n)), p, n, "delta_filter")


def delta2maf(delta: bytes, ref_fasta: bytes, qry_fasta: bytes) -> bytes:
    """`delta2maf` on .delta text and the two FASTA texts (lib/nucmer/mugsy_nucmer.ml:118-124)."""
    p, n = C.c_void_p(), C.c_size_t()
    return _text_out(lib().pmo_delta2maf(delta, len(delta), ref_fasta, len(ref_fasta), qry_fasta, len(qry_fasta), C.byref(p), C.byref(n)), p, n, "delta2maf")
