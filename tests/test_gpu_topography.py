"""Analytic topography evaluated on the device (kgpu_set_topography_function): tiles
activated during the run take their heights from tile_topog_kernel instead of the heights callback.  For the
algebraic functions (xslope, xparab: what the reference's dynamic-tile inputs use) the run must be bit-identical
to the callback run, activated tiles included; for the transcendental ones the heights agree to libm rounding."""
import os

import numpy as np
import pytest

from common import INPUTS, compare_snapshots
from kestrel_b200 import capi
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.run import Simulation
from kestrel_b200.host.settings import RunSet
from kestrel_b200.host.topog import tile_heights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case,kw", [
    ("case_tile_indep_dynamic_20m.txt", dict(tend=8.0, Nout=2)),
    ("case_cap_dilute.txt", dict(tend=6.0, Nout=1)),
    ("case_cap_morpho.txt", dict(tend=4.0, Nout=2)),
])
def test_dynamic_tiles_with_device_topography_are_bit_identical(gpu_lib, case, kw):
    def run(dev):
        rs = read_input_file(os.path.join(INPUTS, case))
        for k, v in kw.items():
            setattr(rs, k, v)
        rs.finalize()
        return Simulation(rs, gpu_lib, device_topography=dev).run()
    a, b = run(False), run(True)
    assert a.infos[-1].ntiles_added > 0, "the case must activate tiles during the run"
    assert [(i.nsteps, i.nrefines, i.ntiles_added, i.t) for i in a.infos] == [(i.nsteps, i.nrefines, i.ntiles_added, i.t) for i in b.infos]
    res = compare_snapshots(a.snapshots[-1], b.snapshots[-1])
    assert all(exact for _, exact in res.values()), res
    for tid in a.snapshots[-1]:
        assert np.array_equal(a.snapshots[-1][tid]["b0"], b.snapshots[-1][tid]["b0"])


@pytest.mark.parametrize("func,params", [
    ("flat", []), ("xslope", [-0.04]), ("yslope", [0.03]), ("xyslope", [-0.08, 0.03]), ("xsinslope", [0.3]), ("xysinslope", [0.2]),
    ("xhump", [0.5, 9.0]), ("xtanh", [2.0, 0.4, 5.0]), ("xparab", [0.002]), ("xyparab", [0.002, 0.001]), ("xbislope", [20.0, 5.0, 3.0]),
    ("x2slopes", [0.5, 0.1, 12.0]), ("usgs", [0.3, 4.0]), ("flume", [28.0, 3.0, 6.0, 1.5, 0.4, 3.0]),
    ("channel power law", [-0.05, 4.0, 2.5]), ("channel trapezium", [-0.05, 5.0, 0.8]), ("xtrislope", [25.0, 10.0, 2.0, 4.0, -6.0, 5.0]),
])
def test_every_topography_function_matches_the_host(gpu_lib, func, params):
    """Heights of a freshly activated tile: device kernel against kestrel_b200/host/topog.py (TopogFuncs.f90)."""
    rs = RunSet(nXtiles=5, nYtiles=4, nXpertile=12, nYpertile=10, Xtilesize=6.0, bcs="periodic", topog_func=func, topog_params=params).finalize()
    p, keep = rs.to_c()
    st = capi.Stepper(gpu_lib, p, keep)
    st.set_topography_function(func, params)
    tid = 8
    nX, nY = rs.nXpertile, rs.nYpertile
    u = np.zeros((nY, nX, 13))
    st.upload_tile(tid, u, b0v=None)          # no heights given: the library evaluates them
    got = st.download_tile(tid)["b0"]
    ref = tile_heights(rs, tid)
    algebraic = func in ("flat", "xslope", "yslope", "xyslope", "xparab", "xyparab")   # no libm call at all
    tol = 32 * 2.2e-16 * max(1.0, float(np.max(np.abs(ref))))
    assert np.max(np.abs(got - ref)) <= tol
    if algebraic:
        # the tile's own vertices (the last row / column belong to the neighbours that were loaded as its ghost tiles:
        # EqualiseTopographicBoundaryData, the same rule on both paths) carry the host's bits
        assert np.array_equal(got[:-1, :-1], ref[:-1, :-1])
    st.close()
