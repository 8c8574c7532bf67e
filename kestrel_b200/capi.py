"""ctypes binding of the C-ABI declared in include/kestrel_gpu.h.

The same binding class drives any shared library that exports that ABI under a
symbol prefix: the product library (prefix ``kgpu_``, built from
kestrel_b200/csrc) and -- from tests/ and bench.py's cpu_baseline leg only -- the
CPU oracle (prefix ``kor_``).  Nothing in this package loads the oracle.

The struct mirrors ``kgpu_params``; field order and types must match the header.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB_PATH = os.path.join(_HERE, "lib", "libkestrel_gpu.so")
if os.environ.get("KGPU_LIB"):  # tuning experiments: another build of the same sources (tools/build_variant.py)
    GPU_LIB_PATH = os.environ["KGPU_LIB"]

# status codes (include/kestrel_gpu.h)
KGPU_OK, KGPU_ERR_ARG, KGPU_ERR_CUDA, KGPU_ERR_HALT_BC, KGPU_ERR_DT, KGPU_ERR_UNSUPPORTED = range(6)

BCS = {"halt": 0, "periodic": 1, "dirichlet": 2, "sponge": 3}
LIMITERS = {"minmod1": 0, "minmod2": 1, "none": 2, "van albada": 3, "albada": 3, "weno": 4}
DRAGS = {"chezy": 0, "coulomb": 1, "voellmy": 2, "pouliquen": 3, "edwards2019": 4, "variable": 5, "manning": 6}
EROSIONS = {"off": 0, "simple": 1, "fluid": 2, "granular": 3, "mixed": 4, "on": 4}
DEPOSITIONS = {"none": 0, "simple": 1, "spearman manning": 2}
ERO_TRANSITIONS = {"smooth": 0, "step": 1, "off": 2}
MORPHO_DAMPS = {"none": 0, "off": 0, "tanh": 1, "rat3": 2}
SWITCHES = {"tanh": 0, "rat3": 1, "cos": 2, "linear": 3, "equal": 4, "0.5": 4, "off": 5, "0": 5,
            "zero": 5, "1": 6, "one": 6, "step": 7}


TOPOG_FUNCS = {"flat": 0, "xslope": 1, "yslope": 2, "xyslope": 3, "xsinslope": 4, "xysinslope": 5, "xhump": 6, "xtanh": 7,
               "xparab": 8, "xyparab": 9, "xbislope": 10, "x2slopes": 11, "usgs": 12, "flume": 13, "channel power law": 14,
               "channel_powerlaw": 14, "channel trapezium": 15, "channel_trapezium": 15, "xtrislope": 16}


class KgpuSource(C.Structure):
    _fields_ = [
        ("x", C.c_double), ("y", C.c_double), ("radius", C.c_double),
        ("num_cells_in_src", C.c_int32), ("n_series", C.c_int32),
        ("time", C.POINTER(C.c_double)), ("flux", C.POINTER(C.c_double)), ("psi", C.POINTER(C.c_double)),
    ]


class KgpuCap(C.Structure):   # kgpu_cap
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("radius", C.c_double), ("height", C.c_double), ("u", C.c_double),
                ("v", C.c_double), ("psi", C.c_double), ("shape", C.c_int32), ("_pad", C.c_int32)]


class KgpuCube(C.Structure):  # kgpu_cube
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("length", C.c_double), ("width", C.c_double), ("height", C.c_double),
                ("u", C.c_double), ("v", C.c_double), ("psi", C.c_double), ("shape", C.c_int32), ("_pad", C.c_int32)]


SHAPES = {"flat": 0, "para": 1, "level": 2}
HEIGHTS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_double))


class KgpuParams(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_int32),
        ("nXpertile", C.c_int32), ("nYpertile", C.c_int32), ("nXtiles", C.c_int32), ("nYtiles", C.c_int32),
        ("isOneD", C.c_int32),
        ("deltaX", C.c_double), ("deltaY", C.c_double), ("xSize", C.c_double), ("ySize", C.c_double),
        ("bcs", C.c_int32), ("_pad0", C.c_int32),
        ("bcsHnval", C.c_double), ("bcsuval", C.c_double), ("bcsvval", C.c_double), ("bcspsival", C.c_double),
        ("geometric_factors", C.c_int32), ("MorphodynamicsOn", C.c_int32),
        ("g", C.c_double), ("rhow", C.c_double), ("rhos", C.c_double), ("gred", C.c_double),
        ("ChezyCo", C.c_double), ("ManningCo", C.c_double), ("CoulombCo", C.c_double),
        ("PouliquenMinSlope", C.c_double), ("PouliquenMaxSlope", C.c_double),
        ("PouliquenIntermediateSlope", C.c_double), ("PouliquenBeta", C.c_double),
        ("Edwards2019betastar", C.c_double), ("Edwards2019kappa", C.c_double), ("Edwards2019Gamma", C.c_double),
        ("VoellmySwitchRate", C.c_double), ("VoellmySwitchValue", C.c_double),
        ("EroRate", C.c_double), ("EroRateGranular", C.c_double), ("CriticalShields", C.c_double),
        ("EroDepth", C.c_double), ("EroCriticalHeight", C.c_double),
        ("BedPorosity", C.c_double), ("maxPack", C.c_double), ("SolidDiameter", C.c_double),
        ("ws0", C.c_double), ("nsettling", C.c_double), ("EddyViscosity", C.c_double),
        ("heightThreshold", C.c_double),
        ("cfl", C.c_double), ("diffusiveTimeScale", C.c_double), ("maxdt", C.c_double),
        ("tstart", C.c_double),
        ("TileBuffer", C.c_int32), ("SpongeLayer", C.c_int32),
        ("SpongeStrength", C.c_double),
        ("limiter", C.c_int32), ("drag", C.c_int32), ("erosion", C.c_int32), ("deposition", C.c_int32),
        ("erosion_transition", C.c_int32), ("morpho_damp", C.c_int32), ("fswitch", C.c_int32),
        ("n_sources", C.c_int32),
        ("sources", C.POINTER(KgpuSource)),
        ("heights", HEIGHTS_FN),
        ("heights_ctx", C.c_void_p),
        ("device", C.c_int32), ("arithmetic", C.c_int32),
        ("comm_rank", C.c_int32), ("comm_size", C.c_int32), ("comm_px", C.c_int32), ("comm_py", C.c_int32),
    ]


class KgpuStepInfo(C.Structure):
    _fields_ = [("t", C.c_double), ("dt_last", C.c_double), ("nsteps", C.c_int64),
                ("nrefines", C.c_int64), ("ntiles_added", C.c_int64)]


class KestrelError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


_dp = C.POINTER(C.c_double)


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class Library:
    """A shared library exporting the kestrel_gpu.h ABI under ``prefix``."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "kgpu_" else C.RTLD_LOCAL)
        f = self._fn
        f("create", C.c_int, [C.POINTER(KgpuParams), C.POINTER(C.c_void_p)])
        f("destroy", C.c_int, [C.c_void_p])
        f("last_error", C.c_char_p, [C.c_void_p])
        f("upload_tile", C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp, C.c_int32])
        f("integrate_to", C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.POINTER(KgpuStepInfo)])
        f("active_tiles", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)])
        f("ghost_tiles", C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)])
        f("download_tile", C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp, _dp, _dp])
        f("upload_domain", C.c_int, [C.c_void_p, _dp, _dp, _dp])
        f("download_domain", C.c_int, [C.c_void_p, _dp, _dp])
        f("version", C.c_char_p, [])
        # optional / library specific
        for name, res, args in [
            ("debug_rhs", C.c_int, [C.c_void_p, C.c_int32, _dp, _dp, _dp]),
            ("download_field", C.c_int, [C.c_void_p, C.c_int32, _dp]),
            ("set_threads", C.c_int, [C.c_void_p, C.c_int]),
            ("launch_count", C.c_int64, [C.c_void_p]),
            ("rhs_timing", C.c_int, [C.c_void_p, _dp, C.POINTER(C.c_int64), C.c_int32]),
            ("stream", C.c_void_p, [C.c_void_p]),
            ("comm_id_bytes", C.c_int, []),
            ("comm_create_id", C.c_int, [C.c_void_p]),
            ("comm_attach", C.c_int, [C.c_void_p, C.c_void_p]),
            ("comm_block", C.c_int, [C.c_void_p] + [C.POINTER(C.c_int32)] * 4),
            ("set_pinned", C.c_int, [C.c_void_p, C.c_int32]),
            ("set_topography_function", C.c_int, [C.c_void_p, C.c_int32, _dp, C.c_int32]),
            ("output_begin", C.c_int, [C.c_void_p, _dp, _dp]),
            ("output_wait", C.c_int, [C.c_void_p]),
            ("morpho_stats", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
            ("set_topography_raster", C.c_int, [C.c_void_p, _dp, C.c_int32, C.c_int32] + [C.c_double] * 6),
            ("load_source_conditions", C.c_int, [C.c_void_p, C.POINTER(KgpuCap), C.c_int32, C.POINTER(KgpuCube), C.c_int32,
                                                 C.POINTER(C.c_int32)]),
            ("debug_sequential_walk", C.c_int, [C.c_void_p, C.c_int32]),
            ("debug_global_walk", C.c_int, [C.c_void_p, C.c_int32]),
            ("debug_morpho_fusion", C.c_int, [C.c_void_p, C.c_int32]),
            ("debug_redist_capacity", C.c_int, [C.c_void_p, C.c_int32]),
        ]:
            try:
                f(name, res, args)
            except AttributeError:
                pass

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.dll, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, name, fn)

    def has(self, name: str) -> bool:
        return hasattr(self, name)


_gpu_lib: Optional[Library] = None


def load_gpu() -> Library:
    """Load the product library.  Raises if it has not been built -- no fallback."""
    global _gpu_lib
    if _gpu_lib is None:
        _gpu_lib = Library(GPU_LIB_PATH, "kgpu_")
    return _gpu_lib


class Stepper:
    """Host-side handle mirroring the calls the Fortran host makes around IntegrateTo
    (TimeStepper.f90:73-113): create -> upload -> integrate_to -> download."""

    def __init__(self, lib: Library, params: "KgpuParams", keepalive: Sequence = ()):
        self.lib = lib
        self.params = params
        self._keep = list(keepalive)
        self.h = C.c_void_p()
        params.struct_bytes = C.sizeof(KgpuParams)
        rc = lib.create(C.byref(params), C.byref(self.h))
        if rc != 0:
            raise KestrelError(rc, "create failed (is a CUDA device visible?)" if lib.prefix == "kgpu_" else "create failed")
        self.nX, self.nY = params.nXpertile, params.nYpertile
        self.nXt, self.nYt = params.nXtiles, params.nYtiles
        if params.comm_size > 1:  # this handle owns one block of the tile grid
            self.nXt, self.nYt = params.nXtiles // params.comm_px, params.nYtiles // params.comm_py
        self.NX, self.NY = self.nX * self.nXt, self.nY * self.nYt
        self.oneD = bool(params.isOneD)

    # -- helpers
    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.last_error(self.h)
            raise KestrelError(rc, msg.decode() if msg else "error")

    def close(self):
        if self.h:
            self.lib.destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state in
    def upload_tile(self, tile_id: int, u13: np.ndarray, b0v=None, btv=None, maxima=None, tfirst=None,
                    contains_source: bool = False):
        u13 = np.ascontiguousarray(u13, dtype=np.float64)
        assert u13.size == 13 * self.nX * self.nY
        self._check(self.lib.upload_tile(self.h, tile_id, _ptr(u13), _ptr(b0v), _ptr(btv), _ptr(maxima),
                                         _ptr(tfirst), int(contains_source)))

    def upload_domain(self, q4: np.ndarray, b0v: np.ndarray, btv: Optional[np.ndarray] = None):
        q4 = np.ascontiguousarray(q4, dtype=np.float64)
        b0v = np.ascontiguousarray(b0v, dtype=np.float64)
        assert q4.size == 4 * self.NX * self.NY
        assert b0v.size == (self.NX + 1) * (1 if self.oneD else self.NY + 1)
        self._check(self.lib.upload_domain(self.h, _ptr(q4), _ptr(b0v), _ptr(btv)))

    # -- the path
    def integrate_to(self, tend: float, max_steps: int = 0) -> KgpuStepInfo:
        info = KgpuStepInfo()
        self._check(self.lib.integrate_to(self.h, float(tend), int(max_steps), C.byref(info)))
        return info

    # -- state out
    def active_tiles(self) -> np.ndarray:
        n = C.c_int32()
        self._check(self.lib.active_tiles(self.h, C.byref(n), None))
        ids = np.zeros(max(n.value, 1), dtype=np.int32)
        self._check(self.lib.active_tiles(self.h, C.byref(n), ids.ctypes.data_as(C.POINTER(C.c_int32))))
        return ids[: n.value]

    def ghost_tiles(self) -> np.ndarray:
        n = C.c_int32()
        self._check(self.lib.ghost_tiles(self.h, C.byref(n), None))
        ids = np.zeros(max(n.value, 1), dtype=np.int32)
        self._check(self.lib.ghost_tiles(self.h, C.byref(n), ids.ctypes.data_as(C.POINTER(C.c_int32))))
        return ids[: n.value]

    def download_tile(self, tile_id: int):
        nX, nY = self.nX, self.nY
        nvy = 1 if self.oneD else nY + 1
        u13 = np.zeros((nY, nX, 13))
        b0v = np.zeros((nY + 1, nX + 1))
        btv = np.zeros((nY + 1, nX + 1))
        maxima = np.zeros((5, 2, nY, nX))
        tfirst = np.zeros((nY, nX))
        self._check(self.lib.download_tile(self.h, tile_id, _ptr(u13), _ptr(b0v), _ptr(btv), _ptr(maxima), _ptr(tfirst)))
        return {"u": u13, "b0": b0v[:nvy], "bt": btv[:nvy], "maxima": maxima, "tfirst": tfirst}

    def download_domain(self, want_bt: bool = False):
        q4 = np.zeros((4, self.NY, self.NX))
        btv = np.zeros(((1 if self.oneD else self.NY + 1), self.NX + 1)) if want_bt else None
        self._check(self.lib.download_domain(self.h, _ptr(q4), _ptr(btv)))
        return (q4, btv) if want_bt else q4

    def set_topography_function(self, name: Optional[str], params: Sequence[float] = ()):
        """Evaluate the analytic topography `name` (Topog function of the input file) on the device for every tile
        activated from now on; None returns to the heights callback."""
        func = -1 if name is None else TOPOG_FUNCS[name.lower()]
        arr = np.ascontiguousarray(list(params), dtype=np.float64)
        self._check(self.lib.set_topography_function(self.h, func, _ptr(arr) if arr.size else None, int(arr.size)))

    def set_topography_raster(self, elev: np.ndarray, origin_x: float, origin_y: float, pixel_w: float, pixel_h: float,
                              centre_e: float = 0.0, centre_n: float = 0.0):
        """Resample the heights of every tile activated from now on from this raster section on the device
        (TileHeightData, dem.f90:260-356).  elev[j, i] = Elev(i + 1, j + 1): x fastest."""
        elev = np.ascontiguousarray(elev, dtype=np.float64)
        ny, nx = elev.shape
        self._check(self.lib.set_topography_raster(self.h, _ptr(elev), nx, ny, origin_x, origin_y, pixel_w, pixel_h, centre_e, centre_n))

    def output_begin(self, q4: np.ndarray, btv: Optional[np.ndarray] = None):
        """Asynchronous output gather: device snapshot now, transfer into q4 (and btv) while the next
        kgpu_integrate_to calls run.  The arrays must stay alive and untouched until output_wait()."""
        assert q4.size == 4 * self.NX * self.NY
        self._check(self.lib.output_begin(self.h, _ptr(q4), _ptr(btv)))

    def output_wait(self):
        self._check(self.lib.output_wait(self.h))

    def load_source_conditions(self, caps: Sequence = (), cubes: Sequence = (), n_sources: int = 0):
        """LoadSourceConditions on the device (kgpu_load_source_conditions): caps / cubes are the host dataclasses of
        kestrel_b200.host.settings.  Returns NumCellsInSrc of the handle's flux sources."""
        ca = (KgpuCap * max(1, len(caps)))()
        for k, c in enumerate(caps):
            ca[k] = KgpuCap(c.x, c.y, c.radius, c.height, c.u, c.v, c.psi, SHAPES[c.shape], 0)
        cu = (KgpuCube * max(1, len(cubes)))()
        for k, c in enumerate(cubes):
            cu[k] = KgpuCube(c.x, c.y, c.length, c.width, c.height, c.u, c.v, c.psi, SHAPES[c.shape], 0)
        counts = (C.c_int32 * max(1, n_sources))()
        self._check(self.lib.load_source_conditions(self.h, ca, len(caps), cu, len(cubes), counts))
        return [int(counts[k]) for k in range(n_sources)]

    def morpho_stats(self):
        """(cells handed to RedistributeGrid, enlargements of its list buffer) since creation."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.morpho_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def debug_rhs(self, substep: int = 1):
        E = np.zeros((4, self.NY, self.NX))
        I = np.zeros((self.NY, self.NX))
        dt = C.c_double()
        self._check(self.lib.debug_rhs(self.h, substep, _ptr(E), _ptr(I), C.byref(dt)))
        return E, I, dt.value

    def assemble(self, fields: Sequence[int] = tuple(range(13))):
        """Gather active tiles into flat (len(fields), NY, NX) arrays; inactive cells are NaN."""
        out = np.full((len(fields), self.NY, self.NX), np.nan)
        for tid in self.active_tiles():
            tx, ty = (tid - 1) % self.nXt, (tid - 1) // self.nXt
            d = self.download_tile(int(tid))
            blk = d["u"][:, :, list(fields)]
            out[:, ty * self.nY:(ty + 1) * self.nY, tx * self.nX:(tx + 1) * self.nX] = np.moveaxis(blk, 2, 0)
        return out
