"""Builds the native pieces in-tree (they travel to the GPU box with the snapshot):
   _lib/libpmnucmer.so  — CUDA kernels + C ABI, sm_100a only
   _lib/libpmn_synth.so — synthetic genome generator (plain C)
   _lib/nucmer          — `nucmer`-argv-compatible CLI shim
   _lib/delta-filter, _lib/delta2maf — argv-compatible front ends of the two post-steps
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PMN_LIB_DIR: an instrumented build (e.g. PMN_NVCC_EXTRA=-DPMN_STITCH_TIMING) goes to its own directory next to _lib
LIB = os.path.join(HERE, os.environ.get("PMN_LIB_DIR", "_lib"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v" if os.environ.get("PMN_PTXAS_V") else "-O3"] + os.environ.get("PMN_NVCC_EXTRA", "").split()


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    cu = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = cu + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    so = os.path.join(LIB, "libpmnucmer.so")
    if force or _stale(so, deps):
        objs = []
        for src in cu:
            obj = os.path.join(LIB, os.path.basename(src)[:-3] + ".o")
            if force or _stale(obj, deps):
                cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
                if verbose:
                    print(" ".join(cmd))
                subprocess.check_call(cmd)
            objs.append(obj)
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "lib64")
        subprocess.check_call(["g++", "-shared", "-o", so] + objs + ["-L", cuda_lib, "-lcudart", "-lpthread", "-Wl,-rpath," + cuda_lib])
    synth = os.path.join(LIB, "libpmn_synth.so")
    src = os.path.join(CSRC, "pmn_synth.c")
    if force or _stale(synth, [src]):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-o", synth, src])
    cli_src = os.path.join(CSRC, "nucmer_main.cpp")
    cli = os.path.join(LIB, "nucmer")
    if os.path.exists(cli_src) and (force or _stale(cli, [cli_src, so])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), cli_src, "-o", cli,
                               "-L", LIB, "-lpmnucmer", "-Wl,-rpath,$ORIGIN"])
    post_src = os.path.join(CSRC, "post_main.cpp")
    post = os.path.join(LIB, "delta-filter")
    if os.path.exists(post_src) and (force or _stale(post, [post_src, so])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(HERE, "..", "include"), post_src, "-o", post,
                               "-L", LIB, "-lpmnucmer", "-Wl,-rpath,$ORIGIN"])
        import shutil
        shutil.copy2(post, os.path.join(LIB, "delta2maf"))      # one binary, the program name decides
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
