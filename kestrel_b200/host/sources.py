"""Initial conditions on the host: LoadSourceConditions (SetSources.f90:47-392).

Rasterises caps / cubes into per-tile ``u(13,nX,nY)`` arrays, counts
NumCellsInSrc and marks containsSource -- the step *before* the path, kept on the
host exactly as in the reference; the result is handed to kgpu_upload_tile.
Quirks reproduced: Q9 (<= R^2 membership), Q10 (1-D flat caps add capu/cappsi to
the derived u/psi; 1-D 'para' momentum without the parabola factor), rho blending
at :159,206,299,351.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .topog import tile_coords, tile_heights

# zero-based indices of u(d,:,:) (main.f90:76-101)
W, HU, HV, HPSI, HN, U, V, PSI, RHO, B0, BT, BX, BY = range(13)


def kahan_sum(terms):
    """Element-wise Kahan sum of a list of arrays (utilities.f90:418-448)."""
    s = np.zeros_like(terms[0])
    c = np.zeros_like(terms[0])
    for x in terms:
        y = x - c
        t = s + y
        c = (t - s) - y
        s = t
    return s


def centre_topography(rs, b0v: np.ndarray, btv: np.ndarray = None):
    """ComputeCellCentredTopographicData (MorphodynamicRHS.f90:308-368) for one tile.
    b0v[(nY+1),(nX+1)] -> b0c, btc, bx, by, each [nY, nX]."""
    if btv is None:
        btv = np.zeros_like(b0v)
    if not rs.isOneD:
        a, b, c, d = b0v[:-1, :-1], b0v[:-1, 1:], b0v[1:, :-1], b0v[1:, 1:]  # (i,j),(i+1,j),(i,j+1),(i+1,j+1)
        ta, tb, tc, td = btv[:-1, :-1], btv[:-1, 1:], btv[1:, :-1], btv[1:, 1:]
        b0c = 0.25 * kahan_sum([a, b, c, d])
        btc = 0.25 * kahan_sum([ta, tb, tc, td])
        bx = 0.5 * rs.deltaXRecip * kahan_sum([b, tb, -a, -ta, d, td, -c, -tc])
        by = 0.5 * rs.deltaYRecip * kahan_sum([c, tc, -a, -ta, d, td, -b, -tb])
    else:
        a, b = b0v[:1, :-1], b0v[:1, 1:]
        ta, tb = btv[:1, :-1], btv[:1, 1:]
        b0c = 0.5 * (a + b)
        btc = 0.5 * (ta + tb)
        bx = rs.deltaXRecip * kahan_sum([b, tb, -a, -ta])
        by = np.zeros_like(bx)
    return b0c, btc, bx, by


def gamma(rs, bx, by):
    if rs.geometric_factors:
        return np.sqrt(1.0 + bx * bx + by * by)
    return np.ones_like(bx)


class Tile:
    def __init__(self, rs, tile_id: int):
        nX, nY = rs.nXpertile, rs.nYpertile
        self.id = tile_id
        self.b0v = tile_heights(rs, tile_id)
        b0c, btc, bx, by = centre_topography(rs, self.b0v)
        u = np.zeros((nY, nX, 13))
        u[..., RHO] = rs.rhow                      # AllocateU, UpdateTiles.f90:243
        u[..., B0], u[..., BT], u[..., BX], u[..., BY] = b0c, btc, bx, by
        u[..., W] = u[..., B0]                     # ActivateTile, UpdateTiles.f90:368
        self.u = u
        self.maxima = np.zeros((5, 2, nY, nX))     # Hnmax, umax, emax, dmax, psimax
        self.tfirst = np.full((nY, nX), -1.0)
        self.contains_source = False


def on_domain_edge(rs, tile_id: int) -> bool:  # Grid.f90:322-335
    i = (tile_id - 1) % rs.nXtiles + 1
    j = (tile_id - 1) // rs.nXtiles + 1
    on = (i == 1 or i == rs.nXtiles)
    return on or (rs.nYtiles > 1 and (j == 1 or j == rs.nYtiles))


def load_source_conditions(rs) -> Dict[int, Tile]:
    """Returns {tile_id: Tile} for the initially active tiles; updates
    rs.sources[k].num_cells_in_src in place."""
    tiles: Dict[int, Tile] = {}
    for s in rs.sources:
        s.num_cells_in_src = 0
    rhow, rhos = rs.rhow, rs.rhos

    def get_tile(k):
        if k not in tiles:
            if on_domain_edge(rs, k) and rs.bcs != "periodic":
                if rs.bcs == "halt":
                    raise RuntimeError("Error: tried to add a tile outside the domain.")  # UpdateTiles.f90:63-65
                return None
            tiles[k] = Tile(rs, k)
        return tiles[k]

    ntx, nty = rs.nXtiles, (1 if rs.isOneD else rs.nYtiles)
    for i in range(1, ntx + 1):
        for j in range(1, nty + 1):
            k = i + (j - 1) * rs.nXtiles
            x, y, _, _ = tile_coords(rs, k)
            X = x[None, :] + 0.0 * y[:, None]
            Y = y[:, None] + 0.0 * x[None, :]
            for cap in rs.caps:
                rho = rhow + (rhos - rhow) * cap.psi
                R2 = (X - cap.x) * (X - cap.x) if rs.isOneD else (X - cap.x) * (X - cap.x) + (Y - cap.y) * (Y - cap.y)
                m = R2 <= cap.radius * cap.radius
                if not m.any():
                    continue
                T = get_tile(k)
                if T is None:
                    continue
                u = T.u
                gam = gamma(rs, u[..., BX], u[..., BY])
                Hn_orig, rho_orig = u[..., HN].copy(), u[..., RHO].copy()
                Hn = np.full_like(gam, cap.height)
                if cap.shape == "flat":
                    u[..., W][m] += (cap.height / gam)[m]
                    u[..., HN][m] += cap.height
                    T.maxima[0, 0][m] += cap.height
                    u[..., HU][m] += rho * cap.height * cap.u
                    if rs.isOneD:
                        T.maxima[4, 0][m] += cap.psi
                        u[..., U][m] += cap.u
                        u[..., PSI][m] += cap.psi
                    else:
                        u[..., HV][m] += rho * cap.height * cap.v
                    u[..., HPSI][m] += cap.psi * cap.height
                elif cap.shape == "para":
                    prof = cap.height * (1.0 - R2 / cap.radius / cap.radius)
                    u[..., W][m] += (prof / gam)[m]
                    u[..., HN][m] += prof[m]
                    T.maxima[0, 0][m] += prof[m]
                    u[..., HU][m] += rho * cap.height * cap.u
                    if rs.isOneD:
                        T.maxima[4, 0][m] += cap.psi
                    else:
                        u[..., HV][m] += rho * cap.height * cap.v
                    u[..., HPSI][m] += (cap.psi * cap.height * (1.0 - R2 / cap.radius / cap.radius))[m]
                elif cap.shape == "level":
                    hp = cap.height - u[..., B0]
                    if rs.isOneD:
                        ok = m & (hp > 0.0)
                        u[..., W][ok] += hp[ok]
                        T.maxima[0, 0][ok] += (hp * gam)[ok]
                        u[..., HN][ok] += (hp * gam)[ok]
                        u[..., HU][ok] += (rho * hp * gam * cap.u)[ok]
                        u[..., HPSI][ok] += (cap.psi * hp * gam)[ok]
                        T.maxima[4, 0][ok] += cap.psi
                    else:
                        Hn = hp * gam
                        ok = m & (Hn > 0.0)
                        u[..., W][ok] += hp[ok]
                        u[..., HN][ok] += Hn[ok]
                        T.maxima[0, 0][ok] += Hn[ok]
                        u[..., HU][ok] += (rho * Hn * cap.u)[ok]
                        u[..., HV][ok] += (rho * Hn * cap.v)[ok]
                        u[..., HPSI][ok] += (cap.psi * Hn)[ok]
                with np.errstate(invalid="ignore", divide="ignore"):
                    newrho = (rho_orig * Hn_orig + rho * Hn) / (Hn_orig + Hn)
                u[..., RHO][m] = newrho[m]
            for cube in rs.cubes:
                rho = rhow + (rhos - rhow) * cube.psi
                m = np.abs(X - cube.x) <= 0.5 * cube.length
                if not rs.isOneD:
                    m = m & (np.abs(Y - cube.y) <= 0.5 * cube.width)
                if not m.any():
                    continue
                T = get_tile(k)
                if T is None:
                    continue
                u = T.u
                gam = gamma(rs, u[..., BX], u[..., BY])
                Hn_orig, rho_orig = u[..., HN].copy(), u[..., RHO].copy()
                Hn = np.full_like(gam, cube.height)
                if cube.shape == "level":
                    hp = cube.height - u[..., B0]
                    Hn = hp * gam
                    ok = m & (Hn > 0.0)
                    u[..., W][ok] += hp[ok]
                    u[..., HN][ok] += Hn[ok]
                    T.maxima[0, 0][ok] += Hn[ok]
                    u[..., HPSI][ok] += (Hn * cube.psi)[ok] if not rs.isOneD else (cube.psi * Hn)[ok]
                else:  # flat
                    u[..., W][m] += (cube.height / gam)[m]
                    u[..., HN][m] += cube.height
                    T.maxima[0, 0][m] += cube.height
                    u[..., HPSI][m] += cube.psi * cube.height
                with np.errstate(invalid="ignore", divide="ignore"):
                    newrho = (rho_orig * Hn_orig + rho * Hn) / (Hn_orig + Hn)
                u[..., RHO][m] = newrho[m]
            for s in rs.sources:
                R2 = (X - s.x) * (X - s.x) if rs.isOneD else (X - s.x) * (X - s.x) + (Y - s.y) * (Y - s.y)
                m = R2 <= s.radius * s.radius
                if not m.any():
                    continue
                T = get_tile(k)
                if T is None:
                    continue
                T.contains_source = True
                s.num_cells_in_src += int(m.sum())
    return tiles
