"""P-GPU == 1-GPU bitwise (skipped unless at least two CUDA devices are visible)."""
import os
import subprocess
import sys

import pytest

from common import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposed_run_is_bitwise_equal(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "run_multigpu.py"), "--tiles", "8", "--per", "32", "--steps", "25"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_decomposed_morphodynamic_run_is_bitwise_equal(world):
    """Strang-split run H M H across ranks: halo exchange of the stage beds, E - D and the centre planes,
    max-allreduce of the refine flags; state AND bed must be bitwise equal to the 1-GPU run."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29520 + world), os.path.join(ROOT, "tests", "run_multigpu.py"), "--tiles", "8", "--per", "32", "--steps", "12",
           "--morpho"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_decomposed_redistribution_is_bitwise_equal(world):
    """RedistributeGrid across ranks (gathered patches, replicated global walk): the thin-layer dam-break
    redistributes excess deposit every step from the 11th on, with the dam fronts sitting on the rank seams."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "run_multigpu.py"), "--tiles", "4", "--per", "32", "--steps", "20",
           "--thin"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "PASS" in r.stdout


SMALL_TILES_2D = ["--set", "nXpertile=10", "--set", "nYpertile=10", "--set", "Xtilesize=10.0", "--set", "Ytilesize=None", "--set", "Nout=2"]
DYNAMIC_CASES = [
    # the reference's inputs with 10-cell tiles, so that the active set grows across the rank seams within tens of steps
    # (the oracle adds 12 / 12 / 11 tiles on these; tests/run_multigpu_dynamic.py requires growth)
    ("case_flux_hydro_2d.txt", SMALL_TILES_2D),                          # flux source sitting on the seams
    ("case_cap_conc_2d.txt", SMALL_TILES_2D + ["--set", "tend=8.0"]),   # released cap with solids
    ("case_flux_hydro.txt", ["--set", "nXpertile=10", "--set", "Xtilesize=10.0", "--set", "tend=30.0", "--set", "Nout=2"]),   # 1-D
    # morphodynamics (Strang step across ranks with dynamic tiles; RedistributeGrid's replicated walk leaves the cells of
    # ghost and inactive tiles alone, as the single-device walk does)
    ("case_cap_morpho_2d.txt", SMALL_TILES_2D + ["--set", "tend=2.0"]),                  # ~4600 redistributed cells, active set across the seam
    ("case_tile_indep_dynamic_20m.txt", ["--set", "tend=2.0", "--set", "Nout=2"]),       # world 2: one rank holds ghost tiles only
    ("case_flux_morpho_2d.txt", SMALL_TILES_2D + ["--set", "tend=5.0"]),
    ("case_cap_morpho.txt", ["--set", "nXpertile=20", "--set", "Xtilesize=20.0", "--set", "tend=5.0", "--set", "Nout=2"]),   # 1-D
    ("case_flux_morpho.txt", ["--set", "nXpertile=10", "--set", "Xtilesize=10.0", "--set", "tend=10.0", "--set", "Nout=2"]),  # 1-D, 2600 refinements
    # an 8 x 8 tile grid the flux source fills within 20 s: dirichlet edges (ghost tiles on the domain edge carry the
    # boundary state, inflow included; 35 of the 36 interior tiles end up active) ...
    ("case_flux_hydro_2d.txt", SMALL_TILES_2D + ["--set", "nXtiles=8", "--set", "nYtiles=8", "--set", "tend=20.0", "--set", "bcs=dirichlet",
                                                 "--set", "bcsHnval=0.02", "--set", "bcsuval=0.1"]),
    # ... and halt edges: every rank stops with KGPU_ERR_HALT_BC at the same step as one device does
    ("case_flux_hydro_2d.txt", SMALL_TILES_2D + ["--set", "nXtiles=8", "--set", "nYtiles=8", "--set", "tend=20.0", "--expect-halt"]),
]


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("case,extra", DYNAMIC_CASES)
def test_decomposed_dynamic_tiles_are_bitwise_equal(world, case, extra):
    """Non-periodic domain, tiles activated by the flow across rank seams (replicated tile table, kgpu_dyn_host.inl):
    active and ghost sets, counters, fields, bed, maxima and heights per tile bitwise equal to the 1-GPU run."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29560 + world), os.path.join(ROOT, "tests", "run_multigpu_dynamic.py"), "--case", case] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "PASS" in r.stdout
