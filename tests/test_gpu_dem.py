"""DEM ingest on the device (kgpu_set_topography_raster): the bicubic raster -> vertex resample of
TileHeightData (dem.f90:260-356) / Bicubic_r (Interp2d.f90:214-292) against a NumPy transcription of the same
statements on a synthetic raster.  Same operations in the same order on both sides: bit-identical."""
import numpy as np
import pytest

from kestrel_b200 import capi
from kestrel_b200.host.settings import RunSet
from kestrel_b200.host.topog import tile_coords

pytestmark = pytest.mark.gpu

# Interp2d.f90:57-74
M = np.array([[1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0],
              [-3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0], [2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0],
              [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0],
              [0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0], [0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0],
              [-3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0],
              [9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1], [-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1],
              [2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0],
              [-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1], [4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1]], dtype=np.float64)


def bicubic_r(x, y, elev):
    """Bicubic_r (Interp2d.f90:214-292); elev[j, i] = z(i + 1, j + 1)."""
    ny, nx = elev.shape
    x0, x1, y0, y1 = int(np.floor(x)), int(np.ceil(x)), int(np.floor(y)), int(np.ceil(y))

    def z(i, j):
        return elev[min(max(j, 1), ny) - 1, min(max(i, 1), nx) - 1]
    dx0, dx1, dy0, dy1 = float(x1 - x0 + 1), float(x1 + 1 - x0), float(y1 - y0 + 1), float(y1 + 1 - y0)
    f = [z(x0, y0), z(x1, y0), z(x0, y1), z(x1, y1),
         (z(x1, y0) - z(x0 - 1, y0)) / dx0, (z(x1 + 1, y0) - z(x0, y0)) / dx1, (z(x1, y1) - z(x0 - 1, y1)) / dx0, (z(x1 + 1, y1) - z(x0, y1)) / dx1,
         (z(x0, y1) - z(x0, y0 - 1)) / dy0, (z(x1, y1) - z(x1, y0 - 1)) / dy0, (z(x0, y1 + 1) - z(x0, y0)) / dy1, (z(x1, y1 + 1) - z(x1, y0)) / dy1,
         ((z(x1, y1) - z(x0 - 1, y1)) - (z(x1, y0 - 1) - z(x0 - 1, y0 - 1))) / dx0 / dy0,
         ((z(x1 + 1, y1) - z(x0, y1)) - (z(x1 + 1, y0 - 1) - z(x0, y0 - 1))) / dx1 / dy0,
         ((z(x1, y1 + 1) - z(x0 - 1, y1 + 1)) - (z(x1, y0) - z(x0 - 1, y0))) / dx0 / dy1,
         ((z(x1 + 1, y1 + 1) - z(x0, y1 + 1)) - (z(x1 + 1, y0) - z(x0, y0))) / dx1 / dy1]
    a = []
    for r in range(16):
        acc = 0.0
        for c in range(16):
            acc = acc + M[r, c] * f[c]
        a.append(acc)
    t = 0.0 if x == x0 else (x - x0) / float(x1 - x0)
    u = 0.0 if y == y0 else (y - y0) / float(y1 - y0)
    ans = 0.0
    for i in range(3, -1, -1):
        ans = t * ans + ((a[i * 4 + 3] * u + a[i * 4 + 2]) * u + a[i * 4 + 1]) * u + a[i * 4 + 0]
    return ans


def tile_height_data(rs, tid, elev, ox, oy, pw, ph, ce, cn):
    """TileHeightData (dem.f90:318-340)."""
    _, _, xv, yv = tile_coords(rs, tid)
    b0 = np.zeros((len(yv), len(xv)))
    for j, Y in enumerate(yv):
        for i, X in enumerate(xv):
            s = 0.0
            for ii in (1, 2):
                for jj in (1, 2):
                    E = ce + X + (float(ii - 1) - 0.5) * rs.deltaX
                    N = cn + Y + (float(jj - 1) - 0.5) * rs.deltaY
                    s = s + bicubic_r((E - ox) / pw + 1, (N - oy) / ph + 1, elev)
            b0[j, i] = 0.25 * s
    return b0


@pytest.mark.parametrize("pixel,integer_aligned", [(5.0, False), (2.5, True)])
def test_raster_resample_matches_the_transcription(gpu_lib, pixel, integer_aligned):
    rs = RunSet(nXtiles=3, nYtiles=3, nXpertile=8, nYpertile=6, Xtilesize=20.0, bcs="periodic", topog_func="flat", topog_params=[]).finalize()
    nx, ny = 64, 48
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    elev = 30.0 + 4.0 * np.sin(0.21 * ii) * np.cos(0.17 * jj) + 0.05 * ii - 0.08 * jj + np.round(3.0 * np.sin(1.3 * ii + 0.7 * jj))
    ce, cn = 500000.0, 4100000.0
    # north-up raster: origin = NW corner, pixel height negative; the domain (60 m x 45 m) sits well inside
    ox = ce - (40.0 if not integer_aligned else 0.5 * pixel * nx)
    oy = cn + (37.0 if not integer_aligned else 0.5 * pixel * ny)
    p, keep = rs.to_c()
    st = capi.Stepper(gpu_lib, p, keep)
    st.set_topography_raster(elev, ox, oy, pixel, -pixel, ce, cn)
    tid = 5
    st.upload_tile(tid, np.zeros((rs.nYpertile, rs.nXpertile, 13)), b0v=None)     # heights from the raster kernel
    got = st.download_tile(tid)["b0"]
    ref = tile_height_data(rs, tid, elev, ox, oy, pixel, -pixel, ce, cn)
    assert np.ptp(ref) > 0.5
    assert np.array_equal(got[:-1, :-1], ref[:-1, :-1]), float(np.max(np.abs(got - ref)))
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    st.close()
