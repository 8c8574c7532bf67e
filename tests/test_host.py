"""Host-side mirror of the reference's setup code: input reader, derived constants,
topography, LoadSourceConditions."""
import math
import os

import numpy as np

from common import INPUTS
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.settings import RunSet
from kestrel_b200.host.sources import centre_topography, kahan_sum, load_source_conditions
from kestrel_b200.host.topog import tile_coords, tile_heights


def test_reader_defaults_and_derived_constants():
    rs = read_input_file(os.path.join(INPUTS, "case_cap_morpho_2d.txt"))
    assert (rs.nXtiles, rs.nYtiles, rs.nXpertile, rs.nYpertile) == (40, 40, 50, 50)
    assert rs.deltaX == 1.0 and rs.xSize == 2000.0 and not rs.isOneD
    assert rs.bcs == "halt" and rs.drag == "variable" and rs.MorphodynamicsOn
    assert rs.cfl == 0.25 and rs.heightThreshold == 1e-5 and rs.TileBuffer == 3
    # Parameters.f90:629-641
    assert rs.gred == (rs.rhos / rs.rhow - 1.0) * rs.g
    R = (rs.gred / 1.2e-6 / 1.2e-6) ** (1.0 / 3.0) * rs.SolidDiameter
    assert rs.CriticalShields == 0.3 / (1.0 + 1.2 * R) + 0.055 * (1.0 - math.exp(-0.02 * R))
    assert rs.diffusiveTimeScale > 1e300  # no eddy viscosity in this input


def test_one_d_detection_and_default_cfl():
    rs = read_input_file(os.path.join(INPUTS, "case_1d_cap_constslope.txt"))
    assert rs.isOneD and rs.cfl == 0.5 and rs.TileBuffer == 1 and not rs.MorphodynamicsOn


def test_tile_coordinates_follow_grid_to_physical():
    """Grid.f90:348-351, UpdateTiles.f90:310-324."""
    rs = RunSet(nXtiles=6, nYtiles=1, nXpertile=200, nYpertile=1, Xtilesize=200.0).finalize()
    x, y, xv, yv = tile_coords(rs, 3)
    assert x[0] == -0.5 * rs.xSize + rs.deltaX * ((3 - 1.0) * 200 + 0.5)
    assert xv[0] == x[0] - 0.5 * rs.deltaX and xv[-1] == x[-1] + 0.5 * rs.deltaX
    assert len(xv) == 201


def test_kahan_sum_matches_scalar_reference():
    rng = np.random.default_rng(1)
    terms = [rng.standard_normal(64) * 10.0 ** rng.integers(-8, 8) for _ in range(8)]
    v = kahan_sum(terms)
    for n in range(64):
        s = c = 0.0
        for t in terms:
            yy = t[n] - c
            tt = s + yy
            c = (tt - s) - yy
            s = tt
        assert s == v[n]


def test_cap_rasterisation_cap_morpho_2d():
    """SetSources.f90:243-303: the cap of tests/Input_cap_morpho_2d.txt sits on a tile corner, so
    four tiles start active; flat cap adds capHeight/gamma to w."""
    rs = read_input_file(os.path.join(INPUTS, "case_cap_morpho_2d.txt"))
    tiles = load_source_conditions(rs)
    assert sorted(tiles) == [860, 861, 900, 901]
    ncell = sum(int((t.u[..., 4] > 0).sum()) for t in tiles.values())
    assert abs(ncell - math.pi * 100) < 25  # radius-10 disc on a 1 m grid
    for t in tiles.values():
        wet = t.u[..., 4] > 0
        gam = np.sqrt(1 + t.u[..., 11] ** 2 + t.u[..., 12] ** 2)
        assert np.allclose((t.u[..., 0] - t.u[..., 9])[wet], (1.0 / gam)[wet], rtol=0, atol=1e-15)
        assert np.array_equal(t.u[..., 3][wet], np.full(wet.sum(), 0.1 * 1.0))
        assert np.array_equal(t.maxima[0, 0][wet], np.ones(wet.sum()))


def test_flux_source_cell_count():
    """NumCellsInSrc counts with <= R^2 (SetSources.f90:367-372, quirk Q9)."""
    rs = read_input_file(os.path.join(INPUTS, "case_flux_hydro_2d.txt"))
    tiles = load_source_conditions(rs)
    x = np.arange(-20, 20) + 0.5
    X, Y = np.meshgrid(x, x)
    assert rs.sources[0].num_cells_in_src == int(((X * X + Y * Y) <= 100.0).sum())
    assert all(t.contains_source for t in tiles.values())


def test_centre_topography_plane_is_exact():
    rs = RunSet(nXtiles=3, nYtiles=3, nXpertile=8, nYpertile=8, Xtilesize=8.0, topog_func="xyslope", topog_params=[0.25, -0.5]).finalize()
    b0v = tile_heights(rs, 5)
    b0c, btc, bx, by = centre_topography(rs, b0v)
    assert np.allclose(bx, 0.25, atol=1e-14) and np.allclose(by, -0.5, atol=1e-14) and np.all(btc == 0)
