"""RunSet mirror: defaults and derived constants of the reference's settings modules.

Follows Parameters.f90:41-77 (defaults), :629-648 (derived constants),
SolverSettings.f90:42-51,189-203, DomainSettings.f90:173-225.  Produces the POD
``kgpu_params`` that crosses the C-ABI.

Deviation (documented): the reference only assigns some defaults when the
matching closure is selected (e.g. Pouliquen slopes only for Pouliquen/Variable
drag, Parameters.f90:522-535) and leaves the rest uninitialised; here every
field always gets its default.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from .. import capi

HUGE = float(np.finfo(np.float64).max)
VISC_W = 1.2e-6  # Parameters.f90:76


@dataclass
class FluxSource:  # type Sources, RunSettings.f90:101-109
    x: float = 0.0
    y: float = 0.0
    radius: float = 0.0
    time: List[float] = field(default_factory=list)
    flux: List[float] = field(default_factory=list)
    psi: List[float] = field(default_factory=list)
    num_cells_in_src: int = 0


@dataclass
class Cap:  # InitConds.f90:254-460
    x: float = 0.0
    y: float = 0.0
    radius: float = 0.0
    height: float = 0.0
    volume: float = 0.0
    psi: float = 0.0
    u: float = 0.0
    v: float = 0.0
    shape: str = "flat"


@dataclass
class Cube:  # InitConds.f90:526-773
    x: float = 0.0
    y: float = 0.0
    length: float = 0.0
    width: float = 0.0
    height: float = 0.0
    psi: float = 0.0
    u: float = 0.0
    v: float = 0.0
    shape: str = "flat"


@dataclass
class RunSet:
    # Domain
    nXtiles: int = 1
    nYtiles: int = 1
    nXpertile: int = 1
    nYpertile: int = 1
    Xtilesize: float = 1.0
    Ytilesize: Optional[float] = None
    bcs: str = "halt"
    bcsHnval: float = 0.0
    bcsuval: float = 0.0
    bcsvval: float = 0.0
    bcspsival: float = 0.0
    # Parameters
    geometric_factors: bool = True
    g: float = 9.81
    rhow: float = 1000.0
    rhos: float = 2000.0
    ChezyCo: float = 0.01
    ManningCo: float = 0.03
    CoulombCo: float = 0.1
    PouliquenMinSlope: float = 0.1
    PouliquenMaxSlope: float = 0.4
    PouliquenIntermediateSlope: float = 0.2
    PouliquenBeta: float = 0.136
    Edwards2019betastar: float = 0.136
    Edwards2019kappa: float = 1.0
    Edwards2019Gamma: float = 0.0
    VoellmySwitchRate: float = 3.0
    VoellmySwitchValue: float = 0.2
    EroRate: float = 0.001
    EroRateGranular: float = 4.0
    EroDepth: float = 1.0
    EroCriticalHeight: float = 0.01
    BedPorosity: float = 0.35
    maxPack: float = 0.65
    SolidDiameter: float = 1e-3
    EddyViscosity: float = 0.0
    ws0: Optional[float] = None  # 'settling speed'
    drag: str = "chezy"
    erosion: str = "mixed"  # default when the key is absent (Parameters.f90:576-580)
    deposition: str = "spearman manning"
    erosion_transition: str = "smooth"
    morpho_damp: str = "tanh"
    fswitch: str = "tanh"
    # Solver
    limiter: str = "minmod2"
    heightThreshold: float = 1e-6
    TileBuffer: int = 1
    cfl: Optional[float] = None
    maxdt: float = HUGE
    tstart: float = 0.0
    tend: float = 1.0
    SpongeStrength: float = 0.2
    # Output
    Nout: int = 1
    out_dir: str = "results/"
    # Topog
    topog_type: str = "function"
    topog_func: str = "flat"
    topog_params: List[float] = field(default_factory=list)
    # Initial conditions
    caps: List[Cap] = field(default_factory=list)
    cubes: List[Cube] = field(default_factory=list)
    sources: List[FluxSource] = field(default_factory=list)
    # library options
    arithmetic: int = 0
    device: int = -1
    comm_rank: int = 0
    comm_size: int = 1
    comm_px: int = 1
    comm_py: int = 1

    # ---- derived (DomainSettings.f90:173-225)
    def finalize(self) -> "RunSet":
        if self.Ytilesize is None:
            self.Ytilesize = self.Xtilesize * float(self.nYpertile) / float(self.nXpertile)
        self.nTiles = self.nXtiles * self.nYtiles
        self.xSize = self.nXtiles * self.Xtilesize
        self.ySize = self.nYtiles * self.Ytilesize
        self.NX = self.nXpertile * self.nXtiles
        self.NY = self.nYpertile * self.nYtiles
        self.isOneD = (self.nYtiles * self.nYpertile == 1)
        self.deltaX = self.Xtilesize / float(self.nXpertile)
        self.deltaY = self.Ytilesize / float(self.nYpertile)
        self.deltaXRecip = 1.0 / self.deltaX
        self.deltaYRecip = 1.0 / self.deltaY
        if self.cfl is None:  # SolverSettings.f90:189-195
            self.cfl = 0.5 if self.isOneD else 0.25
        self.MorphodynamicsOn = self.erosion.lower() != "off"
        self.SpongeLayer = self.bcs == "sponge"
        # Parameters.f90:629-648
        self.gred = (self.rhos / self.rhow - 1.0) * self.g
        self.Rep = math.sqrt(self.g * self.SolidDiameter) * self.SolidDiameter / VISC_W
        R = (self.gred / VISC_W / VISC_W) ** (1.0 / 3.0) * self.SolidDiameter
        if self.ws0 is None:
            self.ws0 = VISC_W / self.SolidDiameter * (math.sqrt(10.36 * 10.36 + 1.048 * R * R * R) - 10.36)
        self.nsettling = (4.7 + 0.41 * self.Rep ** 0.75) / (1.0 + 0.175 * self.Rep ** 0.75)
        self.CriticalShields = 0.3 / (1.0 + 1.2 * R) + 0.055 * (1.0 - math.exp(-0.02 * R))
        self.diffusiveTimeScale = HUGE
        if self.EddyViscosity > 0.0:
            self.diffusiveTimeScale = min(self.deltaX * self.deltaX / self.EddyViscosity,
                                          self.deltaY * self.deltaY / self.EddyViscosity)
        self.DeltaT = (self.tend - self.tstart) / self.Nout  # OutputSettings.f90:167
        return self

    # ---- the POD that crosses the ABI
    def to_c(self, heights_cb: Optional[Callable] = None):
        """Returns (KgpuParams, keepalive list)."""
        p = capi.KgpuParams()
        keep = []
        p.struct_bytes = C.sizeof(capi.KgpuParams)
        p.nXpertile, p.nYpertile, p.nXtiles, p.nYtiles = self.nXpertile, self.nYpertile, self.nXtiles, self.nYtiles
        p.isOneD = int(self.isOneD)
        p.deltaX, p.deltaY, p.xSize, p.ySize = self.deltaX, self.deltaY, self.xSize, self.ySize
        p.bcs = capi.BCS[self.bcs]
        p.bcsHnval, p.bcsuval, p.bcsvval, p.bcspsival = self.bcsHnval, self.bcsuval, self.bcsvval, self.bcspsival
        p.geometric_factors = int(self.geometric_factors)
        p.MorphodynamicsOn = int(self.MorphodynamicsOn)
        for k in ("g", "rhow", "rhos", "gred", "ChezyCo", "ManningCo", "CoulombCo", "PouliquenMinSlope",
                  "PouliquenMaxSlope", "PouliquenIntermediateSlope", "PouliquenBeta", "Edwards2019betastar",
                  "Edwards2019kappa", "Edwards2019Gamma", "VoellmySwitchRate", "VoellmySwitchValue", "EroRate",
                  "EroRateGranular", "CriticalShields", "EroDepth", "EroCriticalHeight", "BedPorosity", "maxPack",
                  "SolidDiameter", "ws0", "nsettling", "EddyViscosity", "heightThreshold", "cfl",
                  "diffusiveTimeScale", "maxdt", "tstart", "SpongeStrength"):
            setattr(p, k, float(getattr(self, k)))
        p.TileBuffer = int(self.TileBuffer)
        p.SpongeLayer = int(self.SpongeLayer)
        p.limiter = capi.LIMITERS[self.limiter.lower()]
        p.drag = capi.DRAGS[self.drag.lower()]
        p.erosion = capi.EROSIONS[self.erosion.lower()]
        p.deposition = capi.DEPOSITIONS[self.deposition.lower()]
        p.erosion_transition = capi.ERO_TRANSITIONS[self.erosion_transition.lower()]
        p.morpho_damp = capi.MORPHO_DAMPS[self.morpho_damp.lower()]
        p.fswitch = capi.SWITCHES[self.fswitch.lower()]
        p.n_sources = len(self.sources)
        if self.sources:
            arr = (capi.KgpuSource * len(self.sources))()
            for k, s in enumerate(self.sources):
                t = (C.c_double * len(s.time))(*s.time)
                f = (C.c_double * len(s.flux))(*s.flux)
                ps = (C.c_double * len(s.psi))(*s.psi)
                keep += [t, f, ps]
                arr[k].x, arr[k].y, arr[k].radius = s.x, s.y, s.radius
                arr[k].num_cells_in_src = s.num_cells_in_src
                arr[k].n_series = len(s.time)
                arr[k].time = C.cast(t, C.POINTER(C.c_double))
                arr[k].flux = C.cast(f, C.POINTER(C.c_double))
                arr[k].psi = C.cast(ps, C.POINTER(C.c_double))
            keep.append(arr)
            p.sources = C.cast(arr, C.POINTER(capi.KgpuSource))
        if heights_cb is not None:
            cb = capi.HEIGHTS_FN(heights_cb)
            keep.append(cb)
            p.heights = cb
        p.device = self.device
        p.arithmetic = self.arithmetic
        p.comm_rank, p.comm_size, p.comm_px, p.comm_py = self.comm_rank, self.comm_size, self.comm_px, self.comm_py
        return p, keep
