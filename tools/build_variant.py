#!/usr/bin/env python
"""Tuning aid: build a variant of the library with extra -D flags on ONE source file, next to the regular one:
   python tools/build_variant.py NAME pop_tracer.cu -DIV_E_GLOBAL -DIV_CH=4
   -> pop2-cesm_b200/csrc/variants/libpop_b200_NAME.so   (git-ignored; travels to the GPU box)
On the box: cp the variant over csrc/libpop_b200.so, run bench.py, restore."""
import importlib.util
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("popbuild", os.path.join(ROOT, "pop2-cesm_b200", "build.py"))
B = importlib.util.module_from_spec(spec)
spec.loader.exec_module(B)


def main():
    name, src, defs = sys.argv[1], sys.argv[2], sys.argv[3:]
    B.build()
    nvcc = B._nvcc()
    inc, libdir, lib = B._nccl_flags()
    flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false", "-rdc=true",
             "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-Xptxas", "-v"] + inc + defs
    vdir = os.path.join(B.CSRC, "variants")
    os.makedirs(vdir, exist_ok=True)
    obj = os.path.join(vdir, src.replace(".cu", "_%s.o" % name))
    r = subprocess.run([nvcc] + flags + ["-dc", os.path.join(B.CSRC, src), "-o", obj], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.exit(r.stdout)
    for line in r.stdout.splitlines():
        if "impvmixt" in line or "spill" in line or "Used" in line:
            pass
    objs = [obj if s == src else os.path.join(B.CSRC, s.replace(".cu", ".o")) for s in B.SOURCES]
    so = os.path.join(vdir, "libpop_b200_%s.so" % name)
    r2 = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", so]
                        + objs + libdir + lib + ["-lcudart"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r2.returncode != 0:
        sys.exit(r2.stdout)
    open(os.path.join(vdir, "ptxas_%s.log" % name), "w").write(r.stdout)
    print(so)


if __name__ == "__main__":
    main()
