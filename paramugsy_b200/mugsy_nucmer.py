"""Host-side mirror of the reference's worker executable lib/nucmer/mugsy_nucmer.ml.

Same option record, same flags, same file contract, same error behaviour; the one line that
differs is the reference's
    Shell.sh "nucmer %s %s -p %s %s" ref_file query_file obname options.nucmer_opts   (mugsy_nucmer.ml:100)
which here calls the B200 library through its C ABI instead of forking MUMmer.

`delta-filter` (mugsy_nucmer.ml:102-105) and `delta2maf` (:118-124) are external programs in the
reference (MUMmer 3.20 / Mugsy).  Here they are calls into the same library (pmn_delta_filter,
pmn_delta2maf: SURVEY.md §8f rows 1 and 2) on the same context, with the same intermediate files
(nucmer.delta, nucmer.filt.delta) so that -debug runs can be compared file by file; a user
post-processor (-delta_pp) is still run from $PATH, and a missing program raises Failure like
Shell.sh would.
"""
import os
import shlex
import shutil
import subprocess
import sys
from dataclasses import dataclass
from typing import Optional

from . import lib


class Failure(Exception):
    """OCaml's `Failure` (mugsy_nucmer.ml:78-81) / a failing Shell.sh."""


@dataclass
class Options:                      # mugsy_nucmer.ml:30-41
    ref_seq: str
    query_seq: str
    maf_out: str
    delta_out: str
    delta_pp: Optional[str] = None
    nucmer_opts: str = ""
    out_dir: str = "/tmp"
    filter: bool = True
    colinear: bool = False
    debug: bool = False
    tmp_dir: str = "/tmp"


def parse_argv(argv) -> Options:
    """mugsy_nucmer.ml:46-94.  Flags: -out_dir -ref_seq -query_seq -maf_out -delta_out -delta_pp
    -nucmer_opts -nofilter -colinear -debug -tmp_dir."""
    v = dict(out_dir="/tmp", ref_seq="", query_seq="", delta_pp=None, nucmer_opts="", filter=True, colinear=False,
             debug=False, maf_out="", delta_out="", tmp_dir="/tmp")
    takes = {"-out_dir": "out_dir", "-ref_seq": "ref_seq", "-query_seq": "query_seq", "-maf_out": "maf_out",
             "-delta_out": "delta_out", "-delta_pp": "delta_pp", "-nucmer_opts": "nucmer_opts", "-tmp_dir": "tmp_dir"}
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in takes:
            if i + 1 >= len(argv):
                raise Failure(f"option '{a}' needs an argument")
            v[takes[a]] = argv[i + 1]; i += 2
        elif a == "-nofilter":
            v["filter"] = False; i += 1
        elif a == "-colinear":
            v["colinear"] = True; i += 1
        elif a == "-debug":
            v["debug"] = True; i += 1
        elif a.startswith("-"):
            raise Failure(f"unknown option '{a}'")
        else:
            i += 1                      # anonymous arguments are collected and ignored (mugsy_nucmer.ml:60)
    if v["ref_seq"] == "" or v["query_seq"] == "":
        raise Failure("Must specify -ref_seq and -query_seq")
    if v["maf_out"] == "" or v["delta_out"] == "":
        raise Failure("Must specify -maf_out and -delta_out")
    v["maf_out"] = v["out_dir"] + "/" + v["maf_out"]          # mugsy_nucmer.ml:85-86
    v["delta_out"] = v["out_dir"] + "/" + v["delta_out"]
    return Options(**v)


def nucmer_opts_to_pmn(opts: str) -> lib.Opts:
    """The string the reference hands to nucmer verbatim (mugsy_nucmer.ml:100), as pmn_opts — read by the library's
    own option table (pmn_opts_parse), the one the `nucmer` shim and the OCaml stub use."""
    try:
        return lib.opts_from_nucmer_string(opts)
    except lib.PmnError as e:
        raise Failure(str(e))


def _sh(options: Options, cmd: str):
    if options.debug:
        print(cmd, file=sys.stderr)
    prog = shlex.split(cmd)[0]
    if shutil.which(prog) is None:
        raise Failure(f"{prog}: command not found (external program of the reference, SURVEY.md §8f)")
    rc = subprocess.call(cmd, shell=True)
    if rc != 0:
        raise Failure(f"command failed with status {rc}: {cmd}")


def nucmer(options: Options, ref_file: str, query_file: str, ctx: Optional[lib.Context] = None) -> str:
    """mugsy_nucmer.ml:96-116: <tmp>/nucmer.delta, optional delta-filter, optional post-processor."""
    obname = f"{options.tmp_dir}/nucmer"
    delta_file = f"{obname}.delta"
    delta_filt_file = f"{obname}.filt.delta"
    own = ctx is None
    c = ctx or lib.Context(int(os.environ.get("PMN_DEVICE", "0")))
    try:
        o = nucmer_opts_to_pmn(options.nucmer_opts)
        rc = lib.lib().pmn_align_pair(c.h, os.fsencode(ref_file), os.fsencode(query_file), o, os.fsencode(delta_file))
        if rc != 0:
            raise Failure(lib.lib().pmn_last_error(None).decode(errors="replace"))
        if options.filter:
            # mugsy_nucmer.ml:102-105: "delta-filter %s %s > %s" with -m when -colinear, else -1
            try:
                with open(delta_file, "rb") as f:
                    filtered = c.delta_filter(f.read(), 2 if options.colinear else 1)
            except lib.PmnError as e:
                raise Failure(str(e))
            with open(delta_filt_file, "wb") as f:
                f.write(filtered)
            delta_file = delta_filt_file
    finally:
        if own:
            c.close()
    if options.delta_pp is not None:
        delta_pp_file = f"{obname}.pp.delta"
        _sh(options, f"{options.delta_pp} < {shlex.quote(delta_file)} > {shlex.quote(delta_pp_file)}")
        delta_file = delta_pp_file
    return delta_file


def generate_maf(options: Options, ctx: Optional[lib.Context] = None):
    """mugsy_nucmer.ml:118-124: "delta2maf %s > %s" on delta_out."""
    own = ctx is None
    c = ctx or lib.Context(int(os.environ.get("PMN_DEVICE", "0")))
    try:
        with open(options.delta_out, "rb") as f:
            delta = f.read()
        rs, qs = c.sequence_from_file(options.ref_seq), c.sequence_from_file(options.query_seq)
        try:
            maf = c.delta2maf(delta, rs, qs)
        finally:
            qs.close(); rs.close()
    except lib.PmnError as e:
        raise Failure(str(e))
    finally:
        if own:
            c.close()
    with open(options.maf_out, "wb") as f:
        f.write(maf)


def run_search(options: Options, ctx=None, maf=True):
    """mugsy_nucmer.ml:127-131."""
    delta_file = nucmer(options, options.ref_seq, options.query_seq, ctx)
    shutil.copyfile(delta_file, options.delta_out)
    if maf:
        generate_maf(options, ctx)


def main(argv=None):
    """mugsy_nucmer.ml:134-140 (including the rm -rf of tmp_dir)."""
    options = parse_argv(sys.argv[1:] if argv is None else argv)
    os.makedirs(options.out_dir, exist_ok=True)
    os.makedirs(options.tmp_dir, exist_ok=True)
    run_search(options)
    shutil.rmtree(options.tmp_dir, ignore_errors=True)


# ---- the callers either side (lib/base/nucmer_task.ml, lib/base/pm_job.ml) ------------------------

def basename(ref_seq: str, query_seq: str) -> str:
    """nucmer_task.ml:10-11."""
    return os.path.basename(ref_seq) + "-" + os.path.basename(query_seq)


def out_paths(tmp_dir: str, sequences):
    """nucmer_task.ml:13-23: keys <bname>-maf / <bname>-delta."""
    m = {}
    for ref_seq, query_seq in sequences:
        b = basename(ref_seq, query_seq)
        m[b + "-maf"] = os.path.join(tmp_dir, b + ".maf")
        m[b + "-delta"] = os.path.join(tmp_dir, b + ".delta")
    return m


def make_commands(searches, tmp_dir: str):
    """nucmer_task.ml:48-59: one mugsy_nucmer command line per pair (note -out_dir gets tmp_dir and
    -tmp_dir gets tmp_dir/<bname>, Appendix A of SURVEY.md)."""
    cmds = []
    for ref_seq, query_seq in searches:
        b = basename(ref_seq, query_seq)
        cmds.append(f"mugsy_nucmer -ref_seq {ref_seq} -query_seq {query_seq} -out_dir {tmp_dir} -tmp_dir {os.path.join(tmp_dir, b)} "
                    f"-maf_out {b}.maf -delta_out {b}.delta")
    return cmds


def searches(genomes):
    """pm_job.ml:43-51: upper-triangle ordered pairs, the earlier genome is the reference."""
    return [(genomes[i], g) for i in range(len(genomes)) for g in genomes[i + 1:]]


def cross(left, right):
    """pm_job.ml:53-57."""
    return [(a, b) for a in left for b in right]


def chunk(n: int, items):
    """job_processor.ml:33-34."""
    return [items[i:i + n] for i in range(0, len(items), n)]


def run_nucmers(searches_, tmp_dir: str, nucmer_chunk: int = 10, ctx=None, filter=True, colinear=False):
    """job_processor.ml:128-154 without the script/queue machinery: every chunk of pairs goes to the
    library's batch entry point (the batch unit of the reference, nucmer_task.ml:6); returns the
    out_paths map.  With `filter` (mugsy_nucmer's default, mugsy_nucmer.ml:54) every pair leaves what one
    mugsy_nucmer process leaves: <bname>.delta = the filtered delta and <bname>.maf (pmn_worker_batch); without it
    (-nofilter) the unfiltered delta and its MAF.  Any failing pair raises Failure (job_processor.ml:72-73 fails the whole node)."""
    import ctypes as C
    os.makedirs(tmp_dir, exist_ok=True)
    own = ctx is None
    c = ctx or lib.Context(int(os.environ.get("PMN_DEVICE", "0")))
    try:
        for part in chunk(nucmer_chunk, list(searches_)):
            outs = [os.path.join(tmp_dir, basename(a, b) + ".delta") for a, b in part]
            arr = lambda xs: (C.c_char_p * len(xs))(*[os.fsencode(x) for x in xs])
            if filter:
                mafs = [os.path.join(tmp_dir, basename(a, b) + ".maf") for a, b in part]
                o = lib.default_opts(post=2 if colinear else 1)
                rc = lib.lib().pmn_worker_batch(c.h, len(part), arr([a for a, _ in part]), arr([b for _, b in part]), arr(outs), arr(mafs), C.byref(o))
            else:
                rc = lib.lib().pmn_align_batch(c.h, len(part), arr([a for a, _ in part]), arr([b for _, b in part]), arr(outs), None)
            if rc != 0:
                raise Failure(lib.lib().pmn_last_error(None).decode(errors="replace"))
            if not filter:
                # -nofilter only skips delta-filter: delta2maf still runs, on the unfiltered delta (mugsy_nucmer.ml:127-131)
                for (a, b), out in zip(part, outs):
                    o = Options(ref_seq=a, query_seq=b, maf_out=out[:-len(".delta")] + ".maf", delta_out=out)
                    generate_maf(o, c)
    finally:
        if own:
            c.close()
    return out_paths(tmp_dir, searches_)


if __name__ == "__main__":
    try:
        main()
    except Failure as e:
        print(f"Fatal error: exception Failure(\"{e}\")", file=sys.stderr)
        sys.exit(2)
