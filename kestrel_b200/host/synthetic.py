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


def thin_dambreak_runset(n_tiles: int, per_tile: int = 32, **over) -> RunSet:
    """Morphodynamic dam-break in a 5-10 cm layer carrying 30 % solids over a 2 cm bed relief: the deposit
    soon exceeds what the thin flow holds, so RedistributeGrid (Redistribute.f90:203) runs from about the
    11th step on -- the workload of the redistribution tests."""
    rs = dambreak_runset(n_tiles, per_tile, morpho=True, topog_params=[0.02], **over)
    L = rs.xSize
    rs.cubes = [Cube(x=0.0, y=0.0, length=L, width=L, height=0.05, psi=0.3, shape="level"),
                Cube(x=-0.25 * L, y=0.0, length=0.5 * L, width=L, height=0.05, psi=0.3, shape="flat")]
    return rs


def dambreak_state(rs: RunSet, block=None):
    """Initial state of the whole domain, or of one block of the tile grid.

    Returns q4[(4, NY, NX)] = (w, rhoHnu, rhoHnv, Hnpsi) and b0v[(NY+1, NX+1)] for the cells
    [i0, i0+NX) x [j0, j0+NY) of the global grid; block = (tx0, ty0, ntx, nty) in tiles (the
    2-D decomposition of kgpu_comm_attach), None = everything.  Same arithmetic as
    LoadSourceConditions on every cell (SetSources.f90:305-355); row blocks keep the
    temporaries small at 16384^2.  Periodic: global vertex NX aliases vertex 0
    (EqualiseTopographicBoundaryData across the wrap), so vertex coordinates wrap."""
    NXg, NYg = rs.NX, rs.NY
    if block is None:
        i0, j0, NX, NY = 0, 0, NXg, NYg
    else:
        tx0, ty0, ntx, nty = block
        i0, j0, NX, NY = tx0 * rs.nXpertile, ty0 * rs.nYpertile, ntx * rs.nXpertile, nty * rs.nYpertile
    xv = -0.5 * rs.xSize + rs.deltaX * ((i0 + np.arange(NX + 1)) % NXg).astype(np.float64)
    yv = -0.5 * rs.ySize + rs.deltaY * ((j0 + np.arange(NY + 1)) % NYg).astype(np.float64)
    q4 = np.zeros((4, NY, NX))
    b0v = np.empty((NY + 1, NX + 1))
    rows = max(1, min(NY, (1 << 22) // max(NX, 1)))
    for r0 in range(0, NY + 1, rows):
        r1 = min(NY + 1, r0 + rows)
        b0v[r0:r1] = topog(rs, xv, yv[r0:r1])
    x = -0.5 * rs.xSize + rs.deltaX * (i0 + np.arange(NX)).astype(np.float64) + 0.5 * rs.deltaX
    yc = -0.5 * rs.ySize + rs.deltaY * (j0 + np.arange(NY)).astype(np.float64) + 0.5 * rs.deltaY
    for r0 in range(0, NY, rows):
        r1 = min(NY, r0 + rows)
        b0c, _, bx, by = centre_topography(rs, b0v[r0:r1 + 1])
        gam = gamma(rs, bx, by)
        y = yc[r0:r1]
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
        q4[0, r0:r1] = w
        q4[3, r0:r1] = hpsi
    return q4, b0v


def decomposition(n: int):
    """(px, py) of the 2-D block decomposition used for n GPUs."""
    return {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(n) or (n, 1)


def rank_block(rs: RunSet, rank: int, px: int, py: int):
    """Tile block (tx0, ty0, ntx, nty) owned by `rank` -- the rule of kgpu_create."""
    ntx, nty = rs.nXtiles // px, rs.nYtiles // py
    return ((rank % px) * ntx, (rank // px) * nty, ntx, nty)
