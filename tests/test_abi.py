"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares.
No compute call is made here (there is no GPU); kgpu_create must fail loudly, not fall back."""
import ctypes
import os
import re

import pytest

from common import ROOT
from kestrel_b200 import capi
from kestrel_b200 import build as kbuild


def declared_symbols(headers=("kestrel_gpu.h", "kestrel_gpu_debug.h")):
    out = set()
    for name in headers:
        hdr = open(os.path.join(ROOT, "include", name)).read()
        hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
        out |= set(re.findall(r"\b(kgpu_[a-z_0-9]+)\s*\(", hdr))
    return sorted(out)


@pytest.fixture(scope="module")
def built_lib():
    return kbuild.build()


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["kgpu_create", "kgpu_destroy", "kgpu_last_error", "kgpu_upload_tile", "kgpu_integrate_to", "kgpu_active_tiles",
              "kgpu_ghost_tiles", "kgpu_download_tile", "kgpu_upload_domain", "kgpu_download_domain", "kgpu_comm_attach"]:
        assert s in syms


def test_public_header_carries_no_test_probes():
    """The drop-in boundary (kestrel_gpu.h) is free of kgpu_debug_* entry points; they live in kestrel_gpu_debug.h."""
    assert not [s for s in declared_symbols(("kestrel_gpu.h",)) if "debug" in s]
    assert any("debug" in s for s in declared_symbols(("kestrel_gpu_debug.h",)))


def test_library_exports_every_declared_symbol(built_lib):
    dll = ctypes.CDLL(built_lib)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, missing


def test_struct_size_matches_header(built_lib):
    """kgpu_create rejects a params struct whose size differs from the C definition."""
    dll = ctypes.CDLL(built_lib)
    p = capi.KgpuParams()
    p.struct_bytes = ctypes.sizeof(capi.KgpuParams) + 8
    h = ctypes.c_void_p()
    dll.kgpu_create.restype = ctypes.c_int
    assert dll.kgpu_create(ctypes.byref(p), ctypes.byref(h)) == capi.KGPU_ERR_ARG


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device kgpu_create reports KGPU_ERR_CUDA; nothing is emulated on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from kestrel_b200.host.synthetic import dambreak_runset
    rs = dambreak_runset(1, 16)
    p, keep = rs.to_c()
    with pytest.raises(capi.KestrelError) as ei:
        capi.Stepper(capi.load_gpu(), p, keep)
    assert ei.value.code == capi.KGPU_ERR_CUDA


def test_product_never_references_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may touch oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "kestrel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                txt = open(os.path.join(dp, f)).read()
                if re.search(r"libkestrel_oracle|kor_[a-z]|oracle/", txt):
                    # capi.py documents the prefix in a docstring only
                    if f == "capi.py" and "kor_" in txt and "libkestrel_oracle" not in txt:
                        continue
                    bad.append(f)
    assert not bad, bad
