"""The contracted-arithmetic variant (params.arithmetic = 1: FMA contraction, shared reciprocals,
max-rate CFL) against the oracle.  Tolerance: the north-star bar, rel-Linf <= 1e-10 per field
after the same step count; active-tile sets and step counts must still match exactly."""
import os

import numpy as np
import pytest

from common import INPUTS, compare_snapshots, domain_stepper, rel_linf, run_input
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.mark.parametrize("ntiles,per,kw", [(2, 64, {}), (3, 50, {}), (2, 32, dict(EddyViscosity=0.05)),
                                            (2, 32, dict(limiter="weno")), (2, 32, dict(geometric_factors=False))])
def test_fast_dambreak(oracle_lib, gpu_lib, ntiles, per, kw):
    rs = dambreak_runset(ntiles, per, **kw)
    q4, b0v = dambreak_state(rs)
    so = domain_stepper(oracle_lib, rs, q4, b0v)
    rs.arithmetic = 1
    sg = domain_stepper(gpu_lib, rs, q4, b0v)
    io, ig = so.integrate_to(1e9, 60), sg.integrate_to(1e9, 60)
    assert (io.nsteps, io.nrefines) == (ig.nsteps, ig.nrefines)
    assert abs(io.t - ig.t) <= 1e-12 * io.t
    qo, qg = so.download_domain(), sg.download_domain()
    for d, name in enumerate(["w", "rhoHnu", "rhoHnv", "Hnpsi"]):
        assert rel_linf(qg[d], qo[d]) <= TOL, (name, rel_linf(qg[d], qo[d]))


def test_fast_dambreak_partial_solids(oracle_lib, gpu_lib):
    """Solids only in the upper cube: CTA tiles of pure water take the no-solids face loop, the
    others the general one, and fronts between them move during the run."""
    rs = dambreak_runset(2, 64)
    rs.cubes[1].psi = 0.2
    q4, b0v = dambreak_state(rs)
    assert (q4[3] == 0).any() and (q4[3] > 0).any()
    so = domain_stepper(oracle_lib, rs, q4, b0v)
    rs.arithmetic = 1
    sg = domain_stepper(gpu_lib, rs, q4, b0v)
    io, ig = so.integrate_to(1e9, 60), sg.integrate_to(1e9, 60)
    assert (io.nsteps, io.nrefines) == (ig.nsteps, ig.nrefines)
    qo, qg = so.download_domain(), sg.download_domain()
    for d, name in enumerate(["w", "rhoHnu", "rhoHnv", "Hnpsi"]):
        assert rel_linf(qg[d], qo[d]) <= TOL, (name, rel_linf(qg[d], qo[d]))


def test_fast_lake_at_rest(gpu_lib):
    path = os.path.join(INPUTS, "case_lake_at_rest_hydro_2d.txt")
    sg = run_input(gpu_lib, path, arithmetic=1)
    Hn0 = sg.snapshots[0][1]["u"][..., 4]
    for snap in sg.snapshots[1:]:
        Hn = snap[1]["u"][..., 4]
        assert np.all(np.char.mod("%.10E", Hn) == np.char.mod("%.10E", Hn0))


@pytest.mark.parametrize("case,kw", [
    ("case_1d_cap_constslope.txt", dict(tend=30.0, Nout=2)),
    ("case_flux_hydro_2d.txt", dict(tend=6.0, Nout=2, nXpertile=20, nYpertile=20, nXtiles=40, nYtiles=40, Xtilesize=20.0, TileBuffer=6)),
    ("case_cap_morpho.txt", dict(tend=2.0, Nout=1)),
])
def test_fast_reference_inputs(oracle_lib, gpu_lib, case, kw):
    path = os.path.join(INPUTS, case)
    sg = run_input(gpu_lib, path, arithmetic=1, **kw)
    so = run_input(oracle_lib, path, **kw)
    assert list(sg.stepper.active_tiles()) == list(so.stepper.active_tiles())
    assert sg.infos[-1].nsteps == so.infos[-1].nsteps
    res = compare_snapshots(sg.snapshots[-1], so.snapshots[-1], fields=[0, 1, 2, 3, 4, 10])
    for name, (err, exact) in res.items():
        assert err <= TOL, f"{name}: {err}"
    v0, vn = sg.volume_rows[0], sg.volume_rows[-1]
    Hn = np.concatenate([t["u"][..., 4].ravel() for t in sg.snapshots[-1].values()])
    assert Hn.min() >= -1e-14


def test_fast_conservation_large(gpu_lib):
    rs = dambreak_runset(16, 128)
    rs.arithmetic = 1
    q4, b0v = dambreak_state(rs)
    from kestrel_b200.host.sources import centre_topography, gamma
    b0c, _, bx, by = centre_topography(rs, b0v)
    g2 = gamma(rs, bx, by) ** 2
    vol0 = float(np.sum((q4[0] - b0c) * g2))
    sg = domain_stepper(gpu_lib, rs, q4, b0v)
    sg.integrate_to(1e9, 40)
    q = sg.download_domain()
    vol1 = float(np.sum((q[0] - b0c) * g2))
    assert abs(vol1 - vol0) / vol0 < 1e-10
    assert np.min(q[0] - b0c) >= -1e-14
