"""Builds csrc/libpop_b200.so (CUDA sm_100a, fp64, -fmad=false) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the dev container as the "does it build"
check; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(CSRC, "libpop_b200.so")
SOURCES = ["pop_core.cu", "pop_grid.cu", "pop_halo.cu", "pop_reduce.cu", "pop_state.cu",
           "pop_tracer.cu", "pop_lwlim.cu", "pop_gm.cu", "pop_momentum.cu", "pop_barotropic.cu", "pop_step.cu", "pop_abi.cu"]
HEADERS = ["pop_ctx.h", "pop_dev.cuh", "pop_state.cuh", "../../include/pop_b200.h"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: the pop_b200 CUDA library cannot be built (no CPU fallback exists)")


def _nccl_flags():
    """Link against NCCL: prefer the copy bundled with torch (the one already in the process when
    bench.py runs under torchrun), fall back to the system library."""
    inc, libdir, lib = [], [], ["-lnccl"]
    try:
        import nvidia.nccl as _n  # torch-bundled wheel
        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        if os.path.exists(os.path.join(base, "include", "nccl.h")):
            inc = ["-I" + os.path.join(base, "include")]
        so2 = os.path.join(base, "lib", "libnccl.so.2")
        if os.path.exists(so2):
            libdir = ["-L" + os.path.join(base, "lib"), "-Xlinker", "-rpath=" + os.path.join(base, "lib")]
            lib = ["-l:libnccl.so.2"]
    except Exception:
        pass
    return inc, libdir, lib


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    nvcc = _nvcc()
    inc, libdir, lib = _nccl_flags()
    objs = []
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
             "-fmad=false", "-rdc=true", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math",
             "-Xptxas", "-v" if verbose else "-O3"] + inc
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(o)
        src = os.path.join(CSRC, s)
        if (not force and os.path.exists(o) and os.path.getmtime(o) > max(
                os.path.getmtime(os.path.join(CSRC, d)) for d in [s] + HEADERS)):
            continue
        procs.append((s, subprocess.Popen([nvcc] + flags + ["-dc", src, "-o", o],
                                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (s, out))
        if verbose:
            sys.stderr.write(out)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-o", SO] + objs + libdir + lib + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
