"""The synthetic dam-break workload C5 of SURVEY.md 8(d) / BASELINE.md 4.

Periodic domain (the only way to have every tile active in the reference,
UpdateTiles.f90:61-69), b0 = 0.2 sin(2 pi x/Lx) sin(2 pi y/Ly) (xySinSlope,
TopogFuncs.f90:486-507), a level cube w = 1.0 over the whole domain plus a flat
cube adding 1 m of depth over the left half (SetSources.f90:305-355), Chezy 0.04,
erosion off, cfl 0.25, MinMod2, height threshold 1e-6, dx = dy = 1 m.
Deterministic functions of (i, j); no RNG.
"""
from __future__ import annotations

import numpy as np

from .settings import Cube, RunSet
from .sources import centre_topography, gamma
from .topog import topog


def dambreak_runset(n_tiles: int, per_tile: int = 128, morpho: bool = False, **over) -> RunSet:
    rs = RunSet(nXtiles=n_tiles, nYtiles=n_tiles, nXpertile=per_tile, nYpertile=per_tile,
                Xtilesize=float(per_tile), bcs="periodic", drag="chezy", ChezyCo=0.04,
                erosion="off", EddyViscosity=0.0, cfl=0.25, heightThreshold=1e-6, limiter="minmod2",
                topog_func="xysinslope", topog_params=[0.2], tend=1.0e9, Nout=1)
    if morpho:  # second run of SURVEY 8(d): the C3 parameter block
        rs.drag, rs.erosion = "variable", "mixed"
        rs.PouliquenMinSlope, rs.PouliquenMaxSlope, rs.PouliquenBeta = 0.1, 0.4, 0.126
        rs.EroRate, rs.EroRateGranular, rs.EroDepth = 1e-3, 0.1, 5.0
        rs.EroCriticalHeight, rs.heightThreshold = 0.01, 1e-5
    for k, v in over.items():
        setattr(rs, k, v)
    rs.finalize()
    L = rs.xSize
    conc = 0.1 if morpho else 0.0
    rs.cubes = [Cube(x=0.0, y=0.0, length=L, width=L, height=1.0, psi=conc, shape="level"),
                Cube(x=-0.25 * L, y=0.0, length=0.5 * L, width=L, height=1.0, psi=conc, shape="flat")]
    return rs


def dambreak_state(rs: RunSet):
    """Flat-domain initial state: q4[(4, NY, NX)] = (w, rhoHnu, rhoHnv, Hnpsi) and b0v[(NY+1, NX+1)].
    Same arithmetic as LoadSourceConditions on every cell (SetSources.f90:305-355), evaluated on the
    whole domain at once (row blocks keep the temporaries small at 16384^2)."""
    NX, NY = rs.NX, rs.NY
    xv = -0.5 * rs.xSize + rs.deltaX * np.arange(NX + 1, dtype=np.float64)
    yv = -0.5 * rs.ySize + rs.deltaY * np.arange(NY + 1, dtype=np.float64)
    q4 = np.zeros((4, NY, NX))
    b0v = np.empty((NY + 1, NX + 1))
    rows = max(1, min(NY, (1 << 22) // max(NX, 1)))
    for j0 in range(0, NY + 1, rows):
        j1 = min(NY + 1, j0 + rows)
        b0v[j0:j1] = topog(rs, xv, yv[j0:j1])
    # periodic: vertex NX aliases vertex 0 (EqualiseTopographicBoundaryData across the wrap)
    b0v[:, NX] = b0v[:, 0]
    b0v[NY, :] = b0v[0, :]
    x = xv[:-1] + 0.5 * rs.deltaX
    for j0 in range(0, NY, rows):
        j1 = min(NY, j0 + rows)
        b0c, _, bx, by = centre_topography(rs, b0v[j0:j1 + 1])
        gam = gamma(rs, bx, by)
        y = yv[j0:j1] + 0.5 * rs.deltaY
        w = b0c.copy()
        hpsi = np.zeros_like(w)
        X = x[None, :] + 0.0 * y[:, None]
        Y = y[:, None] + 0.0 * x[None, :]
        for cube in rs.cubes:
            m = (np.abs(X - cube.x) <= 0.5 * cube.length) & (np.abs(Y - cube.y) <= 0.5 * cube.width)
            if cube.shape == "level":
                hp = cube.height - b0c
                Hn = hp * gam
                ok = m & (Hn > 0.0)
                w[ok] += hp[ok]
                hpsi[ok] += (Hn * cube.psi)[ok]
            else:
                w[m] += (cube.height / gam)[m]
                hpsi[m] += cube.psi * cube.height
        q4[0, j0:j1] = w
        q4[3, j0:j1] = hpsi
    return q4, b0v
