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
