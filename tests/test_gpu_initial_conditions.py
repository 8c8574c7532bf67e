"""LoadSourceConditions on the device (kgpu_load_source_conditions) against the host rasteriser
(kestrel_b200/host/sources.py, SetSources.f90:47-392) followed by kgpu_upload_tile: same active and ghost tiles, the
same state, maxima, NumCellsInSrc bit for bit -- and the same run afterwards (which also checks containsSource and
the seed of the first tile-activation scan)."""
import os

import numpy as np
import pytest

from common import INPUTS, compare_snapshots
from kestrel_b200 import capi
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.run import Simulation
from kestrel_b200.host.settings import Cap, Cube, FluxSource, RunSet
from kestrel_b200.host.topog import make_heights_callback

pytestmark = pytest.mark.gpu


def device_ic(gpu_lib, rs):
    """A stepper whose initial state was rasterised by the library; returns (stepper, NumCellsInSrc)."""
    for s in rs.sources:
        s.num_cells_in_src = 1          # placeholder: the library counts
    p, keep = rs.to_c(make_heights_callback(rs))
    st = capi.Stepper(gpu_lib, p, keep)
    counts = st.load_source_conditions(rs.caps, rs.cubes, len(rs.sources))
    return st, counts


def check_same_initial_state(gpu_lib, rs_factory, steps_tend):
    rs_h = rs_factory()
    host = Simulation(rs_h, gpu_lib)                     # host rasteriser + kgpu_upload_tile
    rs_d = rs_factory()
    st, counts = device_ic(gpu_lib, rs_d)
    assert counts == [s.num_cells_in_src for s in rs_h.sources]
    assert list(st.active_tiles()) == list(host.stepper.active_tiles())
    assert sorted(st.ghost_tiles()) == sorted(host.stepper.ghost_tiles())
    for tid in st.active_tiles():
        a, b = st.download_tile(int(tid)), host.stepper.download_tile(int(tid))
        for d in range(13):
            assert np.array_equal(a["u"][..., d], b["u"][..., d]), (tid, d)
        assert np.array_equal(a["tfirst"], b["tfirst"]) and np.array_equal(a["b0"], b["b0"])
        if rs_h.bcs == "periodic":
            # the host rasteriser takes the depth added to Hnmax from ITS copy of the tile's vertices; across the periodic
            # seam the library holds one copy of each vertex (vertex NX is vertex 0), so sin(+pi) vs sin(-pi) of the wrapped
            # topographies shows up in the last bit of gamma in the tile's last row / column
            assert np.allclose(a["maxima"], b["maxima"], rtol=1e-14, atol=0.0)
        else:
            assert np.array_equal(a["maxima"], b["maxima"])
    ia, ib = st.integrate_to(steps_tend), host.stepper.integrate_to(steps_tend)
    assert (ia.t, ia.nsteps, ia.nrefines, ia.ntiles_added) == (ib.t, ib.nsteps, ib.nrefines, ib.ntiles_added)
    assert list(st.active_tiles()) == list(host.stepper.active_tiles())
    sa = {int(t): st.download_tile(int(t)) for t in st.active_tiles()}
    sb = {int(t): host.stepper.download_tile(int(t)) for t in host.stepper.active_tiles()}
    for name, (err, exact) in compare_snapshots(sa, sb).items():
        assert exact, (name, err)
    for t in sa:
        if rs_h.bcs == "periodic":   # value planes only: whether a later step beats an initial Hnmax that differs in its last bit decides the time plane
            assert np.allclose(sa[t]["maxima"][:, 0], sb[t]["maxima"][:, 0], rtol=1e-14, atol=0.0)
        else:
            assert np.array_equal(sa[t]["maxima"], sb[t]["maxima"])
    st.close()


@pytest.mark.parametrize("case,tend", [
    ("case_1d_cap_constslope.txt", 5.0), ("case_cap_morpho_2d.txt", 0.3), ("case_flux_hydro_2d.txt", 3.0), ("case_cap_conc.txt", 4.0),
    ("case_flux_single_pt.txt", 1.0), ("case_lake_at_rest_hydro_2d.txt", 0.5), ("case_tile_indep_dynamic_20m.txt", 1.0),
])
def test_reference_inputs(gpu_lib, case, tend):
    def make():
        rs = read_input_file(os.path.join(INPUTS, case))
        rs.finalize()
        return rs
    check_same_initial_state(gpu_lib, make, tend)


@pytest.mark.parametrize("oned", [False, True])
def test_every_shape_and_quirk(gpu_lib, oned):
    """All cap shapes (flat, para, level) and cube shapes (flat, level), overlapping, with velocities and solids, 1-D and
    2-D (quirks Q9, Q10), two flux sources, on sloping ground with geometric factors."""
    def make():
        rs = RunSet(nXtiles=9, nYtiles=1 if oned else 9, nXpertile=12, nYpertile=1 if oned else 12, Xtilesize=12.0, bcs="halt",
                    drag="chezy", ChezyCo=0.02, erosion="off", topog_func="xslope" if oned else "xyslope",
                    topog_params=[-0.1] if oned else [-0.1, 0.05], tend=2.0, Nout=1, TileBuffer=2)
        rs.caps = [Cap(x=-6.0, y=0.0, radius=7.5, height=0.8, psi=0.1, u=0.5, v=-0.2, shape="flat"),
                   Cap(x=8.0, y=5.0, radius=9.0, height=1.2, psi=0.05, u=-0.3, v=0.4, shape="para"),
                   Cap(x=2.0, y=-9.0, radius=6.0, height=0.6, psi=0.2, u=0.1, v=0.1, shape="level")]
        rs.cubes = [Cube(x=-10.0, y=8.0, length=9.0, width=7.0, height=0.3, psi=0.15, shape="flat"),
                    Cube(x=12.0, y=-6.0, length=10.0, width=12.0, height=-0.2, psi=0.02, shape="level")]
        rs.sources = [FluxSource(x=0.0, y=0.0, radius=3.0, time=[0.0, 10.0], flux=[2.0, 2.0], psi=[0.0, 0.1]),
                      FluxSource(x=-14.0, y=-12.0, radius=2.0, time=[0.5], flux=[1.0], psi=[0.05])]
        return rs.finalize()
    check_same_initial_state(gpu_lib, make, 1.0)


def test_halt_boundary(gpu_lib):
    """A cap on an edge tile with bcs = halt: AddTile's fatal error (UpdateTiles.f90:63-65) as a status code."""
    rs = RunSet(nXtiles=3, nYtiles=3, nXpertile=4, nYpertile=4, Xtilesize=4.0, bcs="halt", topog_func="flat", topog_params=[])
    rs.caps = [Cap(x=-5.0, y=-5.0, radius=1.0, height=1.0, shape="flat")]
    rs.finalize()
    p, keep = rs.to_c(make_heights_callback(rs))
    st = capi.Stepper(gpu_lib, p, keep)
    with pytest.raises(capi.KestrelError) as ei:
        st.load_source_conditions(rs.caps, rs.cubes, 0)
    assert ei.value.code == capi.KGPU_ERR_HALT_BC
    st.close()
