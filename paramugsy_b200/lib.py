"""ctypes binding of libpmnucmer.so (include/pmnucmer.h).

The extension is built in-tree by paramugsy_b200/build.py (nvcc, sm_100a).  There is no
CPU fallback anywhere: if the library is missing it is built, if it cannot be loaded or no
B200 is visible every computing call raises.
"""
import os as _os
# one hardware work queue per stream (the driver reads this when the CUDA context is created; default 8 makes the
# streams of a scheduler's workers alias): set before anything in this process touches the GPU
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class PmnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpmnucmer error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [("minmatch", C.c_int32), ("mincluster", C.c_int32), ("maxgap", C.c_int32),
                ("diagdiff", C.c_int32), ("diagfactor", C.c_double), ("breaklen", C.c_int32),
                ("do_forward", C.c_int32), ("do_reverse", C.c_int32), ("do_extend", C.c_int32),
                ("do_optimize", C.c_int32), ("do_simplify", C.c_int32), ("keep_stages", C.c_int32), ("post", C.c_int32)]


class NucmerArgs(C.Structure):
    """pmn_nucmer_args: what the `nucmer` command line of lib/nucmer/mugsy_nucmer.ml:100 says."""
    _fields_ = [("opts", Opts), ("prefix", C.c_char_p), ("ref", C.c_char_p), ("qry", C.c_char_p),
                ("device", C.c_int32), ("help", C.c_int32), ("version", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("ref_bases", C.c_int64), ("qry_bases", C.c_int64), ("anchors", C.c_int64),
                ("clusters", C.c_int64), ("cluster_matches", C.c_int64), ("alignments", C.c_int64),
                ("aligned_ref_bases", C.c_int64), ("dp_cells", C.c_int64), ("dp_jobs", C.c_int64),
                ("sa_rounds", C.c_int32), ("kmer_bits", C.c_int32),
                ("ms_index", C.c_float), ("ms_seed", C.c_float), ("ms_cluster", C.c_float),
                ("ms_extend", C.c_float), ("ms_total", C.c_float), ("ms_seed_kernel", C.c_float),
                ("ms_wave1", C.c_float), ("ms_stitch", C.c_float), ("kernel_launches", C.c_int64),
                ("wave1_cells", C.c_int64), ("wall_ms_index", C.c_float), ("wall_ms_align", C.c_float),
                ("wall_ms_text", C.c_float), ("wall_ms_post", C.c_float),
                ("seed_lookups", C.c_int64), ("arena_bytes", C.c_int64), ("seed_probes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/pmnucmer.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pmn_default_opts", "pmn_nucmer_parse_argv", "pmn_opts_parse", "pmn_ctx_create", "pmn_ctx_destroy", "pmn_last_error", "pmn_device_count", "pmn_alloc_count", "pmn_pinned_pool_stats",
    "pmn_ctx_stream", "pmn_ctx_counters", "pmn_ctx_sync_count", "pmn_sched_sync_count", "pmn_measure_int32_peak",
    "pmn_seq_from_fasta", "pmn_seq_from_file", "pmn_seq_free", "pmn_seq_bases", "pmn_seq_records",
    "pmn_index_build", "pmn_index_free", "pmn_index_image", "pmn_index_image_bytes", "pmn_index_alloc", "pmn_index_adopt", "pmn_align", "pmn_seed_part", "pmn_align_anchors", "pmn_result_delta", "pmn_result_stats",
    "pmn_result_free", "pmn_align_pair", "pmn_align_batch", "pmn_index_size", "pmn_index_copy_sa",
    "pmn_result_n_anchors", "pmn_result_copy_anchors", "pmn_result_n_clusters",
    "pmn_result_n_cluster_matches", "pmn_result_copy_clusters", "pmn_result_n_alignments",
    "pmn_result_n_deltas", "pmn_result_copy_alignments",
    "pmn_sched_create", "pmn_sched_destroy", "pmn_sched_workers", "pmn_sched_ctx", "pmn_sched_counters",
    "pmn_sched_align_fasta", "pmn_sched_align_seqs", "pmn_sched_align_indexed", "pmn_sched_align_files",
    "pmn_multi_plan", "pmn_multi_create", "pmn_multi_destroy", "pmn_multi_devices", "pmn_multi_sched", "pmn_multi_align_fasta",
    "pmn_multi_align_files", "pmn_multi_align_large",
    "pmn_delta_filter", "pmn_delta2maf", "pmn_free_text", "pmn_result_filtered", "pmn_result_maf", "pmn_worker_batch",
]


def lib_path():
    return os.path.join(_HERE, os.environ.get("PMN_LIB_DIR", "_lib"), "libpmnucmer.so")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            from . import build
            build.build()
        L = C.CDLL(lib_path())
        vp, cp, i64, i32p, i64p = C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.pmn_default_opts.argtypes = [C.POINTER(Opts)]
        L.pmn_nucmer_parse_argv.argtypes = [C.c_int, C.POINTER(cp), C.POINTER(NucmerArgs)]
        L.pmn_opts_parse.argtypes = [cp, C.POINTER(Opts)]
        L.pmn_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.pmn_ctx_destroy.argtypes = [vp]
        L.pmn_alloc_count.restype = i64
        L.pmn_pinned_pool_stats.argtypes = [i64p]
        L.pmn_last_error.argtypes = [vp]; L.pmn_last_error.restype = cp
        L.pmn_ctx_stream.argtypes = [vp]; L.pmn_ctx_stream.restype = vp
        L.pmn_ctx_counters.argtypes = [vp, i64p]
        L.pmn_ctx_sync_count.argtypes = [vp]; L.pmn_ctx_sync_count.restype = i64
        L.pmn_sched_sync_count.argtypes = [vp]; L.pmn_sched_sync_count.restype = i64
        L.pmn_measure_int32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.pmn_seq_from_fasta.argtypes = [vp, cp, C.c_size_t, C.POINTER(vp)]
        L.pmn_seq_from_file.argtypes = [vp, cp, C.POINTER(vp)]
        L.pmn_seq_free.argtypes = [vp]
        L.pmn_seq_bases.argtypes = [vp]; L.pmn_seq_bases.restype = i64
        L.pmn_seq_records.argtypes = [vp]
        L.pmn_index_build.argtypes = [vp, vp, C.POINTER(vp)]
        L.pmn_index_free.argtypes = [vp]
        L.pmn_index_image.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.pmn_index_image_bytes.argtypes = [i64]; L.pmn_index_image_bytes.restype = C.c_size_t
        L.pmn_index_alloc.argtypes = [vp, vp, C.POINTER(vp)]
        L.pmn_index_adopt.argtypes = [vp]
        L.pmn_align.argtypes = [vp, vp, vp, C.POINTER(Opts), cp, cp, C.POINTER(vp)]
        L.pmn_seed_part.argtypes = [vp, vp, vp, C.POINTER(Opts), C.c_int, C.c_int, C.POINTER(vp), i64p]
        L.pmn_align_anchors.argtypes = [vp, vp, vp, C.POINTER(Opts), vp, i64, cp, cp, C.POINTER(vp)]
        L.pmn_result_delta.argtypes = [vp, C.POINTER(C.c_size_t)]; L.pmn_result_delta.restype = vp
        L.pmn_result_stats.argtypes = [vp, C.POINTER(Stats)]
        L.pmn_result_filtered.argtypes = [vp, C.POINTER(C.c_size_t)]; L.pmn_result_filtered.restype = vp
        L.pmn_result_maf.argtypes = [vp, C.POINTER(C.c_size_t)]; L.pmn_result_maf.restype = vp
        L.pmn_result_free.argtypes = [vp]
        L.pmn_align_pair.argtypes = [vp, cp, cp, C.POINTER(Opts), cp]
        L.pmn_align_batch.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(Opts)]
        L.pmn_worker_batch.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(Opts)]
        L.pmn_index_size.argtypes = [vp]; L.pmn_index_size.restype = i64
        L.pmn_index_copy_sa.argtypes = [vp, vp, vp]
        for name in ("pmn_result_n_anchors", "pmn_result_n_clusters", "pmn_result_n_cluster_matches",
                     "pmn_result_n_alignments", "pmn_result_n_deltas"):
            f = getattr(L, name); f.argtypes = [vp]; f.restype = i64
        L.pmn_result_copy_anchors.argtypes = [vp, vp]
        L.pmn_result_copy_clusters.argtypes = [vp, vp, vp, vp]
        L.pmn_result_copy_alignments.argtypes = [vp, vp, vp, vp]
        i32a = C.POINTER(C.c_int32)
        L.pmn_sched_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.pmn_sched_destroy.argtypes = [vp]
        L.pmn_sched_workers.argtypes = [vp]
        L.pmn_sched_ctx.argtypes = [vp, C.c_int]; L.pmn_sched_ctx.restype = vp
        L.pmn_sched_counters.argtypes = [vp, i64p]
        L.pmn_sched_align_fasta.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(C.c_size_t), C.POINTER(cp), C.c_int, i32a, i32a, C.POINTER(Opts), C.POINTER(vp)]
        L.pmn_sched_align_seqs.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(cp), C.c_int, i32a, i32a, C.POINTER(Opts), C.POINTER(vp)]
        L.pmn_sched_align_indexed.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(cp), C.c_int, i32a, i32a, C.POINTER(Opts), C.POINTER(vp)]
        L.pmn_sched_align_files.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(Opts)]
        L.pmn_multi_plan.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_size_t), C.c_int, i32a, i32a, i32a]
        L.pmn_multi_create.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(vp)]
        L.pmn_multi_destroy.argtypes = [vp]
        L.pmn_multi_devices.argtypes = [vp]
        L.pmn_multi_sched.argtypes = [vp, C.c_int]; L.pmn_multi_sched.restype = vp
        L.pmn_multi_align_fasta.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(C.c_size_t), C.POINTER(cp), C.c_int, i32a, i32a, C.POINTER(Opts), C.POINTER(vp)]
        L.pmn_multi_align_files.argtypes = [vp, C.c_int, C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(cp), C.POINTER(Opts)]
        L.pmn_multi_align_large.argtypes = [vp, cp, C.c_size_t, cp, C.c_size_t, C.POINTER(Opts), cp, cp, C.POINTER(vp), C.POINTER(C.c_double)]
        L.pmn_delta_filter.argtypes = [vp, cp, C.c_size_t, C.c_int, C.c_double, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.pmn_delta2maf.argtypes = [vp, cp, C.c_size_t, vp, vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.pmn_free_text.argtypes = [vp]
        _LIB = L
    return _LIB


def alloc_count() -> int:
    return lib().pmn_alloc_count()


def pinned_pool_stats():
    out = (C.c_int64 * 3)()
    lib().pmn_pinned_pool_stats(out)
    return {"idle_bytes": out[0], "accounted_bytes": out[1], "idle_budget_bytes": out[2]}


def index_image_bytes(n_bases: int) -> int:
    return lib().pmn_index_image_bytes(n_bases)


def default_opts(**kw):
    o = Opts()
    lib().pmn_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown nucmer option {k!r}")
        setattr(o, k, v)
    return o


def opts_from_nucmer_string(nucmer_opts: str) -> Opts:
    """The free-form option string the reference appends to the nucmer command line (mugsy_nucmer.ml:100), read by the
    library's own option table (pmn_opts_parse); raises PmnError(PMN_E_ARG) on an unknown or unsupported option."""
    o = Opts()
    _check(lib().pmn_opts_parse(os.fsencode(nucmer_opts), C.byref(o)))
    return o


def nucmer_parse_argv(argv):
    """argv (without the program name) -> (Opts, prefix, ref, qry, device or None, help, version)."""
    arr = (C.c_char_p * len(argv))(*[os.fsencode(a) for a in argv])
    a = NucmerArgs()
    _check(lib().pmn_nucmer_parse_argv(len(argv), arr, C.byref(a)))
    dec = lambda b: os.fsdecode(b) if b is not None else None
    o = Opts.from_buffer_copy(a.opts)
    return o, dec(a.prefix), dec(a.ref), dec(a.qry), (a.device if a.device >= 0 else None), bool(a.help), bool(a.version)


def _check(rc):
    if rc != 0:
        raise PmnError(rc, lib().pmn_last_error(None).decode(errors="replace"))


class Context:
    """One per (process, GPU): owns the stream and all scratch memory."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().pmn_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if getattr(self, "h", None) and not getattr(self, "borrowed", False):
            lib().pmn_ctx_destroy(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def stream(self) -> int:
        """cudaStream_t of this context as an integer (wrap with torch.cuda.ExternalStream)."""
        return lib().pmn_ctx_stream(self.h) or 0

    def counters(self):
        out = (C.c_int64 * 4)()
        lib().pmn_ctx_counters(self.h, out)
        return {"launches": out[0], "h2d_bytes": out[1], "d2h_bytes": out[2], "pairs": out[3], "syncs": lib().pmn_ctx_sync_count(self.h)}

    def int32_peak(self):
        g, mhz = C.c_double(), C.c_double()
        _check(lib().pmn_measure_int32_peak(self.h, C.byref(g), C.byref(mhz)))
        return g.value, mhz.value

    def _text(self, rc, p, n):
        _check(rc)
        try:
            return C.string_at(p.value, n.value) if n.value else b""
        finally:
            lib().pmn_free_text(p)

    def delta_filter(self, delta: bytes, mode: int = 1, maxolap: float = 75.0) -> bytes:
        """`delta-filter -1` (mode 1) / `-m` (mode 2) on .delta text (lib/nucmer/mugsy_nucmer.ml:102-105)."""
        p, n = C.c_void_p(), C.c_size_t()
        return self._text(lib().pmn_delta_filter(self.h, delta, len(delta), mode, maxolap, C.byref(p), C.byref(n)), p, n)

    def delta2maf(self, delta: bytes, ref: "Sequence", qry: "Sequence") -> bytes:
        """`delta2maf` on .delta text and the two packed genomes it was computed from (mugsy_nucmer.ml:118-124)."""
        p, n = C.c_void_p(), C.c_size_t()
        return self._text(lib().pmn_delta2maf(self.h, delta, len(delta), ref.h, qry.h, C.byref(p), C.byref(n)), p, n)

    def sequence(self, fasta: bytes):
        return Sequence(self, fasta=fasta)

    def sequence_from_file(self, path: str):
        return Sequence(self, path=path)


class Scheduler:
    """W worker threads sharing one GPU (pmn_sched): the in-process form of the reference's
    run_nucmers fan-out (lib/base/job_processor.ml:128-154)."""

    def __init__(self, device=0, workers=8):
        self.h = C.c_void_p()
        _check(lib().pmn_sched_create(device, workers, C.byref(self.h)))
        self.device, self.workers = device, workers

    def close(self):
        if getattr(self, "h", None):
            lib().pmn_sched_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def counters(self):
        out = (C.c_int64 * 4)()
        lib().pmn_sched_counters(self.h, out)
        return {"launches": out[0], "h2d_bytes": out[1], "d2h_bytes": out[2], "pairs": out[3], "syncs": lib().pmn_sched_sync_count(self.h)}

    def context(self, k=0):
        """Worker k's context as a borrowed Context (do not close it)."""
        c = Context.__new__(Context)
        c.h = C.c_void_p(lib().pmn_sched_ctx(self.h, k)); c.device = self.device; c.borrowed = True
        return c

    @staticmethod
    def _pairs(pairs):
        n = len(pairs)
        return n, (C.c_int32 * n)(*[p[0] for p in pairs]), (C.c_int32 * n)(*[p[1] for p in pairs])

    def align_fasta(self, fastas, pairs, names=None, opts=None, **kw):
        """fastas: list of FASTA bytes in host memory; pairs: [(ref index, qry index)] -> [Result]."""
        o = opts if opts is not None else default_opts(**kw)
        g = len(fastas)
        # an entry may also be (address, length) of FASTA text the caller keeps in (pinned) host memory
        addr = [C.cast(C.c_char_p(f), C.c_void_p).value if isinstance(f, (bytes, bytearray)) else int(f[0]) for f in fastas]
        fa = C.cast((C.c_void_p * g)(*addr), C.POINTER(C.c_char_p))
        nb = (C.c_size_t * g)(*[len(f) if isinstance(f, (bytes, bytearray)) else int(f[1]) for f in fastas])
        nm = (C.c_char_p * g)(*[os.fsencode(x) for x in names]) if names else None
        n, r, q = self._pairs(pairs)
        out = (C.c_void_p * n)()
        _check(lib().pmn_sched_align_fasta(self.h, g, fa, nb, nm, n, r, q, C.byref(o), out))
        return [Result(C.c_void_p(h)) for h in out]

    def align_seqs(self, seqs, pairs, names=None, opts=None, indexes=None, **kw):
        """seqs: list of Sequence objects resident on this GPU; indexes: optional list (None entries allowed)
        of Index objects the caller already holds for some of them."""
        o = opts if opts is not None else default_opts(**kw)
        g = len(seqs)
        sh = (C.c_void_p * g)(*[(s.h if s is not None else None) for s in seqs])      # genomes no pair names may be None
        nm = (C.c_char_p * g)(*[os.fsencode(x) for x in names]) if names else None
        n, r, q = self._pairs(pairs)
        out = (C.c_void_p * n)()
        if indexes is None:
            _check(lib().pmn_sched_align_seqs(self.h, g, sh, nm, n, r, q, C.byref(o), out))
        else:
            ih = (C.c_void_p * g)(*[(ix.h if ix is not None else None) for ix in indexes])
            _check(lib().pmn_sched_align_indexed(self.h, g, sh, ih, nm, n, r, q, C.byref(o), out))
        return [Result(C.c_void_p(h)) for h in out]

    def align_files(self, refs, qrys, outs, opts=None, **kw):
        o = opts if opts is not None else default_opts(**kw)
        arr = lambda xs: (C.c_char_p * len(xs))(*[os.fsencode(x) for x in xs])
        _check(lib().pmn_sched_align_files(self.h, len(refs), arr(refs), arr(qrys), arr(outs), C.byref(o)))


def multi_plan(n_devices, pairs, genome_bytes=None):
    """pmn_multi_plan: the device of every pair [(ref, qry)] (pure host arithmetic, no GPU needed)."""
    n = len(pairs)
    ng = (max(max(p) for p in pairs) + 1) if pairs else 0
    if genome_bytes is not None:
        ng = len(genome_bytes)
    r = (C.c_int32 * n)(*[p[0] for p in pairs]); q = (C.c_int32 * n)(*[p[1] for p in pairs]); out = (C.c_int32 * n)()
    nb = (C.c_size_t * ng)(*genome_bytes) if genome_bytes is not None else None
    _check(lib().pmn_multi_plan(n_devices, ng, nb, n, r, q, out))
    return list(out)


class Multi:
    """Several GPUs of one box driven by ONE process (pmn_multi): one scheduler per device."""

    def __init__(self, devices, workers=8):
        devices = list(devices)
        self.h = C.c_void_p()
        _check(lib().pmn_multi_create((C.c_int * len(devices))(*devices), len(devices), workers, C.byref(self.h)))
        self.devices = devices

    def close(self):
        if getattr(self, "h", None):
            lib().pmn_multi_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def align_fasta(self, fastas, pairs, names=None, opts=None, **kw):
        o = opts if opts is not None else default_opts(**kw)
        g = len(fastas)
        fa = (C.c_char_p * g)(*fastas)
        nb = (C.c_size_t * g)(*[len(f) for f in fastas])
        nm = (C.c_char_p * g)(*[os.fsencode(x) for x in names]) if names else None
        n, r, q = Scheduler._pairs(pairs)
        out = (C.c_void_p * n)()
        _check(lib().pmn_multi_align_fasta(self.h, g, fa, nb, nm, n, r, q, C.byref(o), out))
        return [Result(C.c_void_p(h)) for h in out]

    def align_files(self, refs, qrys, outs, mafs=None, opts=None, **kw):
        o = opts if opts is not None else default_opts(**kw)
        arr = lambda xs: (C.c_char_p * len(xs))(*[os.fsencode(x) for x in xs])
        _check(lib().pmn_multi_align_files(self.h, len(refs), arr(refs), arr(qrys), arr(outs), arr(mafs) if mafs else None, C.byref(o)))

    def align_large(self, ref_fasta: bytes, qry_fasta: bytes, ref_path="ref.fa", qry_path="qry.fa", opts=None, **kw):
        """-> (Result, [ms pack+index, ms seeding, ms gather, ms clustering+extension+text])."""
        o = opts if opts is not None else default_opts(**kw)
        r = C.c_void_p(); ms = (C.c_double * 4)()
        _check(lib().pmn_multi_align_large(self.h, ref_fasta, len(ref_fasta), qry_fasta, len(qry_fasta), C.byref(o),
                                          os.fsencode(ref_path), os.fsencode(qry_path), C.byref(r), ms))
        return Result(r), list(ms)


class Sequence:
    def __init__(self, ctx, fasta=None, path=None):
        self.ctx = ctx
        self.h = C.c_void_p()
        if fasta is not None:
            _check(lib().pmn_seq_from_fasta(ctx.h, fasta, len(fasta), C.byref(self.h)))
        else:
            _check(lib().pmn_seq_from_file(ctx.h, os.fsencode(path), C.byref(self.h)))

    @classmethod
    def from_address(cls, ctx, address: int, nbytes: int):
        """FASTA text the caller keeps in (pinned) host memory, given by address and length."""
        self = cls.__new__(cls)
        self.ctx = ctx
        self.h = C.c_void_p()
        _check(lib().pmn_seq_from_fasta(ctx.h, C.cast(C.c_void_p(address), C.c_char_p), nbytes, C.byref(self.h)))
        return self

    @property
    def bases(self):
        return lib().pmn_seq_bases(self.h)

    @property
    def records(self):
        return lib().pmn_seq_records(self.h)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().pmn_seq_free(self.h)
        self.h = None

    def index(self, empty=False):
        return Index(self, empty=empty)


class _DeviceBytes:
    """A range of HBM as a uint8 array for torch.as_tensor (zero copy, __cuda_array_interface__)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class Index:
    def __init__(self, seq, empty=False):
        """Builds the index of `seq`; with empty=True only allocates its image, to be filled by a
        collective (Index.image) and validated with adopt()."""
        self.seq = seq           # keeps the sequence alive
        self.ctx = seq.ctx
        self.h = C.c_void_p()
        if empty:
            _check(lib().pmn_index_alloc(self.ctx.h, seq.h, C.byref(self.h)))
        else:
            _check(lib().pmn_index_build(self.ctx.h, seq.h, C.byref(self.h)))

    def image(self):
        """(device pointer, bytes) of the index image in HBM."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().pmn_index_image(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def image_tensor(self):
        """The image as a torch uint8 CUDA tensor aliasing the index memory (for torch.distributed)."""
        import torch
        p, n = self.image()
        return torch.as_tensor(_DeviceBytes(p, n), device=torch.device("cuda", self.ctx.device))

    def adopt(self):
        _check(lib().pmn_index_adopt(self.h))

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            lib().pmn_index_free(self.h)
        self.h = None

    def suffix_array(self):
        n = lib().pmn_index_size(self.h)
        sa = np.empty(n, np.int32); lcp = np.empty(n, np.int32)
        _check(lib().pmn_index_copy_sa(self.h, sa.ctypes.data, lcp.ctypes.data))
        return sa, lcp

    def align(self, qry, opts=None, ref_path="ref.fa", qry_path="qry.fa", **kw):
        o = opts if opts is not None else default_opts(**kw)
        r = C.c_void_p()
        _check(lib().pmn_align(self.ctx.h, self.h, qry.h, C.byref(o), os.fsencode(ref_path), os.fsencode(qry_path), C.byref(r)))
        return Result(r)


    def seed_part(self, qry, part, nparts, opts=None, **kw):
        """Anchors of query-position part `part` of `nparts` as (device pointer, count); they live in the
        context's scratch until its next call."""
        o = opts if opts is not None else default_opts(**kw)
        p, n = C.c_void_p(), C.c_int64()
        _check(lib().pmn_seed_part(self.ctx.h, self.h, qry.h, C.byref(o), part, nparts, C.byref(p), C.byref(n)))
        return p.value or 0, n.value

    def seed_part_tensor(self, qry, part, nparts, opts=None, **kw):
        """The same as an (n, 4) int32 CUDA tensor (a copy, safe to keep)."""
        import torch
        p, n = self.seed_part(qry, part, nparts, opts, **kw)
        dev = torch.device("cuda", self.ctx.device)
        if n == 0:
            return torch.empty((0, 4), dtype=torch.int32, device=dev)
        return torch.as_tensor(_DeviceBytes(p, 16 * n), device=dev).view(torch.int32).view(n, 4).clone()

    def align_anchors(self, qry, anchors, opts=None, ref_path="ref.fa", qry_path="qry.fa", **kw):
        """Clustering, extension and .delta from a given anchor list ((n, 4) int32 CUDA tensor)."""
        o = opts if opts is not None else default_opts(**kw)
        r = C.c_void_p()
        a = anchors.contiguous()
        _check(lib().pmn_align_anchors(self.ctx.h, self.h, qry.h, C.byref(o), C.c_void_p(a.data_ptr() if a.numel() else 0), a.shape[0],
                                       os.fsencode(ref_path), os.fsencode(qry_path), C.byref(r)))
        return Result(r)


class Result:
    def __init__(self, h):
        self.h = h

    def close(self):
        if getattr(self, "h", None) and _LIB is not None:      # _LIB is gone at interpreter shutdown
            _LIB.pmn_result_free(self.h)
        self.h = None

    __del__ = close

    @property
    def delta(self) -> bytes:
        n = C.c_size_t()
        p = lib().pmn_result_delta(self.h, C.byref(n))
        return C.string_at(p, n.value)

    @property
    def filtered(self) -> bytes:
        """`delta-filter` of .delta (options post=1: -1, post=2: -m); empty without the option."""
        n = C.c_size_t()
        p = lib().pmn_result_filtered(self.h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    @property
    def maf(self) -> bytes:
        """`delta2maf` of the filtered delta (option post); empty without the option."""
        n = C.c_size_t()
        p = lib().pmn_result_maf(self.h, C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    @property
    def stats_raw(self) -> Stats:
        s = Stats()
        lib().pmn_result_stats(self.h, C.byref(s))
        return s

    @property
    def stats(self):
        return self.stats_raw.as_dict()

    def anchors(self):
        n = lib().pmn_result_n_anchors(self.h)
        a = np.empty((n, 4), np.int32)
        if n:
            _check(lib().pmn_result_copy_anchors(self.h, a.ctypes.data))
        return a

    def clusters(self):
        k = lib().pmn_result_n_clusters(self.h); m = lib().pmn_result_n_cluster_matches(self.h)
        ms = np.empty((m, 3), np.int32); off = np.zeros(k + 1, np.int32); tag = np.empty(k, np.int32)
        if k:
            _check(lib().pmn_result_copy_clusters(self.h, ms.ctypes.data, off.ctypes.data, tag.ctypes.data))
        return ms, off, tag

    def alignments(self):
        a = lib().pmn_result_n_alignments(self.h); d = lib().pmn_result_n_deltas(self.h)
        rows = np.empty((a, 10), np.int64); doff = np.zeros(a + 1, np.int64); dl = np.empty(d, np.int64)
        if a:
            _check(lib().pmn_result_copy_alignments(self.h, rows.ctypes.data, doff.ctypes.data, dl.ctypes.data))
        return rows, doff, dl
