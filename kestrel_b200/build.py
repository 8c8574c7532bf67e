"""In-tree build of libkestrel_gpu.so for sm_100a (nvcc cross-compiles without a GPU).

The built library lives in kestrel_b200/lib/ (git-ignored, shipped to the GPU box by
gpurun).  -fmad=false keeps every fp64 operation individually rounded, which is what
makes the CUDA path bit-comparable with the reference's IEEE arithmetic.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libkestrel_gpu.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(HERE, "..", "include", "kestrel_gpu.h"))
    return out


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in deps())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libkestrel_gpu.so.  Files named *_fast.cu hold the
    contracted-arithmetic kernels and are the only ones compiled with -fmad=true."""
    if not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    common = [NVCC, "-ccbin", HOSTCXX, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "-I", os.path.join(HERE, "..", "include")]
    if verbose:
        common += ["-Xptxas", "-v"]
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        fmad = "-fmad=true" if src.endswith("_fast.cu") else "-fmad=false"
        procs.append((src, subprocess.Popen(common + [fmad, "-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = ""
    for src, pr in procs:
        out, _ = pr.communicate()
        log += out
        if pr.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    r = subprocess.run([NVCC, "-ccbin", HOSTCXX, "-shared", "-o", LIB] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if verbose:
        print(log)
    return LIB


HOSTSRC = os.path.join(HERE, "host_cpp")
HOSTBIN = os.path.join(HERE, "bin", "kestrel_gpu_run")


def build_host(force: bool = False) -> str:
    """The C++ host driver above the C-ABI (kestrel_b200/host_cpp): g++ only, links libkestrel_gpu.so
    through an $ORIGIN-relative rpath so the in-tree binary runs wherever the snapshot lands."""
    srcs = [os.path.join(HOSTSRC, f) for f in sorted(os.listdir(HOSTSRC)) if f.endswith(".cpp")]
    dep = [os.path.join(HOSTSRC, f) for f in os.listdir(HOSTSRC)] + [os.path.join(HERE, "..", "include", "kestrel_gpu.h"), LIB]
    if not force and os.path.exists(HOSTBIN) and all(os.path.getmtime(d) <= os.path.getmtime(HOSTBIN) for d in dep):
        return HOSTBIN
    os.makedirs(os.path.dirname(HOSTBIN), exist_ok=True)
    cmd = [HOSTCXX, "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I", os.path.join(HERE, "..", "include"), "-I", HOSTSRC] + srcs + \
          ["-L", LIBDIR, "-lkestrel_gpu", "-Wl,-rpath,$ORIGIN/../lib", "-o", HOSTBIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("host driver build failed")
    return HOSTBIN


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
    print(build_host(force=True))
