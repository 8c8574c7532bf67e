#!/usr/bin/env python
"""Build libkestrel_gpu.so with extra -D defines into kestrel_b200/lib/variants/<name>/ for A/B runs
(select it with KGPU_LIB=<path>).  usage: build_variant.py <name> [-DFOO=1 ...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kestrel_b200 import build as kb  # noqa: E402

name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(kb.LIBDIR, "variants", name)
obj = os.path.join(kb.HERE, "build", "variants", name)
os.makedirs(out, exist_ok=True)
os.makedirs(obj, exist_ok=True)
common = [kb.NVCC, "-ccbin", kb.HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "-Xcompiler", "-O2", "-I", os.path.join(ROOT, "include"), "-Xptxas", "-v"] + defs
procs, objs = [], []
for src in kb.sources():
    o = os.path.join(obj, os.path.basename(src)[:-3] + ".o")
    fmad = "-fmad=true" if src.endswith("_fast.cu") else "-fmad=false"
    procs.append(subprocess.Popen(common + [fmad, "-c", src, "-o", o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    objs.append(o)
for pr in procs:
    log, _ = pr.communicate()
    if pr.returncode:
        sys.exit(log)
    lines = log.splitlines()
    for i, l in enumerate(lines):
        if "hydro_stage_kernelILi32E" in l and "ELb0ELb0ELi1E" in l and "Function properties" in l:
            print(l.split("_ZN4kgpu18")[1][:60], "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
lib = os.path.join(out, "libkestrel_gpu.so")
subprocess.check_call([kb.NVCC, "-ccbin", kb.HOSTCXX, "-shared", "-o", lib] + objs + ["-ldl"])
print(lib)
