"""Analytic topographies and tile coordinates (host side of the ABI).

TopogFuncs.f90 (the functions), Grid.f90:339-353 (GridToPhysical) and
UpdateTiles.f90:288-325 (cell / vertex coordinates of a tile).  This is what the
heights callback evaluates for ``Type = Function`` inputs; DEM / SRTM rasters are
out of scope (GDAL).
Arrays returned are b0[j, i] (i fastest), matching the ABI's vertex layout.
"""
from __future__ import annotations

import math

import numpy as np

PI = 3.141592653589793238462643383279502884


def tile_coords(rs, tile_id: int):
    """x(i), y(j) at cell centres and x_vertex, y_vertex of a tile (1-based id)."""
    gi = (tile_id - 1) % rs.nXtiles + 1
    gj = (tile_id - 1) // rs.nXtiles + 1
    ii = np.arange(1, rs.nXpertile + 1, dtype=np.float64)
    jj = np.arange(1, rs.nYpertile + 1, dtype=np.float64)
    x = -0.5 * rs.xSize + rs.deltaX * ((gi - 1.0) * rs.nXpertile + (ii - 0.5))
    y = -0.5 * rs.ySize + rs.deltaY * ((gj - 1.0) * rs.nYpertile + (jj - 0.5))
    xv = np.concatenate([x - 0.5 * rs.deltaX, [x[-1] + 0.5 * rs.deltaX]])
    yv = np.concatenate([y - 0.5 * rs.deltaY, [y[-1] + 0.5 * rs.deltaY]])
    return x, y, xv, yv


def topog(rs, xv: np.ndarray, yv: np.ndarray) -> np.ndarray:
    """b0 at vertices; returns array [len(yv), len(xv)]."""
    name = rs.topog_func.lower()
    p = list(rs.topog_params)
    X = xv[None, :] + 0.0 * yv[:, None]
    Y = yv[:, None] + 0.0 * xv[None, :]
    if name == "flat":
        return np.zeros_like(X)
    if name == "xslope":
        return p[0] * X
    if name == "yslope":
        return p[0] * Y
    if name == "xyslope":
        return p[0] * X + p[1] * Y
    if name == "xsinslope":
        Lx = rs.Xtilesize * rs.nXtiles
        return p[0] * np.sin(X * (2.0 * PI / Lx))
    if name == "xysinslope":
        Lx = rs.Xtilesize * rs.nXtiles
        Ly = rs.Ytilesize * rs.nYtiles
        return p[0] * np.sin(X * (2.0 * PI / Lx)) * np.sin(Y * (2.0 * PI / Ly))
    if name == "xhump":
        A, L = p[0], p[1]
        return np.where((X > -L) & (X < L), 0.5 * A * (1.0 + np.cos(PI * X / L)), 0.0)
    if name == "xtanh":
        x0, A, L = p[0], p[1], p[2]
        return A * (1.0 + np.tanh((X - x0) / L))
    if name == "xparab":
        return p[0] * X * X
    if name == "xyparab":
        return p[0] * X * X + p[1] * Y * Y
    if name == "xbislope":
        phi1, phi2, lam = p[0] * PI / 180.0, p[1] * PI / 180.0, p[2]
        a1, a2 = math.tan(phi1), math.tan(phi2)
        return -0.5 * (a1 + a2) * X + 0.5 * (a1 - a2) * lam * np.log(np.cosh(X / lam))
    if name == "x2slopes":
        alpha, beta, R = p[0], p[1], p[2]
        sa, sb = math.sqrt(1.0 + alpha * alpha), math.sqrt(1.0 + beta * beta)
        xc0 = (sa - sb) * R / (alpha - beta)
        zc0 = (alpha * sb - beta * sa) * R / (alpha - beta)
        x1 = xc0 - alpha * R / sa
        x2 = xc0 - beta * R / sb
        arc = zc0 - np.sqrt(np.maximum(R * R - (X - xc0) * (X - xc0), 0.0))
        return np.where(X < x1, -alpha * X, np.where(X > x2, -beta * X, arc))
    if name in ("usgs", "flume"):  # TopogFuncs.f90:145-242: two slopes joined by a cosh arc, tanh side walls
        if name == "usgs":
            theta0, theta1, xwall, wallW, wallH, sigma = 31.0, 2.4, 8.5, 2.0, p[0], p[1]
        else:
            theta0, theta1, xwall, wallW, wallH, sigma = p[0], p[1], p[2], p[3], p[4], p[5]
        alpha = 8.5 / (math.asinh(-math.tan(4.0 * PI / 180.0)) - math.asinh(-math.tan(theta0 * PI / 180.0)))
        xc0 = -alpha * math.asinh(-math.tan(theta0 * PI / 180.0))
        zc0 = -alpha * math.cosh((-xc0) / alpha)
        x1 = xc0 + alpha * math.asinh(-math.tan(theta1 * PI / 180.0))
        up = -math.tan(theta0 * PI / 180.0) * X
        down = zc0 + alpha * math.cosh((x1 - xc0) / alpha) - math.tan(theta1 * PI / 180.0) * (X - x1)
        with np.errstate(over="ignore"):
            arc = zc0 + alpha * np.cosh((X - xc0) / alpha)
        b = np.where(X < 0.0, up, np.where(X > x1, down, arc))
        walls = 0.5 * wallH * (np.tanh(sigma * (Y - 0.5 * wallW)) - np.tanh(sigma * (Y - 1.5 * wallW))
                               + np.tanh(sigma * (Y + 1.5 * wallW)) - np.tanh(sigma * (Y + 0.5 * wallW)))
        return np.where(X < xwall, b + walls, b)
    if name in ("channel power law", "channel_powerlaw"):  # TopogFuncs.f90:255-276
        slope, W, alpha = p[0], p[1], p[2]
        costheta = math.cos(math.atan(slope))
        return slope * X + costheta * (np.abs(Y) / W) ** alpha
    if name in ("channel trapezium", "channel_trapezium"):  # TopogFuncs.f90:288-308
        slope, W, Sb = p[0], p[1], p[2]
        costheta = math.cos(math.atan(slope))
        return slope * X + costheta * np.maximum(0.0, Sb * (np.abs(Y) - 0.5 * W))
    if name == "xtrislope":  # TopogFuncs.f90:346-388
        phi1, phi2, phi3 = p[0] * PI / 180.0, p[1] * PI / 180.0, p[2] * PI / 180.0
        lam, x1, x2 = p[3], p[4], p[5]
        s1, s2, s3 = math.tan(phi1), math.tan(phi2), math.tan(phi3)
        c2 = (x1 - 0.5 * lam) * 0.5 * (s1 - s2)
        c3 = (x1 + 0.5 * lam) * 0.5 * (s1 - s2) + c2
        c4 = (x2 - 0.5 * lam) * 0.5 * (s2 - s3) + c3
        c5 = (x2 + 0.5 * lam) * 0.5 * (s2 - s3) + c4
        A12, A23 = 0.5 * (s2 - s1) * lam / PI, 0.5 * (s3 - s2) * lam / PI
        return np.where(X < x1 - 0.5 * lam, s1 * X,
               np.where(X < x1 + 0.5 * lam, A12 * np.sin((X - x1) * PI / lam - 0.5 * PI) + 0.5 * (s1 + s2) * X + c2,
               np.where(X < x2 - 0.5 * lam, s2 * X + c3,
               np.where(X < x2 + 0.5 * lam, A23 * np.sin((X - x2) * PI / lam - 0.5 * PI) + 0.5 * (s2 + s3) * X + c4,
                        s3 * X + c5))))
    raise ValueError(f"topography function '{rs.topog_func}' is not available (raster DEMs are out of scope)")


def tile_heights(rs, tile_id: int) -> np.ndarray:
    """GetHeights for Type = Function (dem.f90:360-415): b0[(nY+1), (nX+1)]."""
    _, _, xv, yv = tile_coords(rs, tile_id)
    return np.ascontiguousarray(topog(rs, xv, yv))


def make_heights_callback(rs):
    """A kgpu_heights_fn closure evaluating tile_heights."""
    n = (rs.nXpertile + 1) * (rs.nYpertile + 1)

    def cb(ctx, tile_id, out):
        try:
            b = tile_heights(rs, int(tile_id)).ravel()
            dst = np.ctypeslib.as_array(out, shape=(n,))
            dst[:] = b
            return 0
        except Exception:  # never raise across the ABI
            return 1

    return cb
