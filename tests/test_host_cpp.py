"""The C++ host driver (kestrel_b200/host_cpp -> kestrel_b200/bin/kestrel_gpu_run) against the
Python host mirror: same input file -> same files.

CPU part (--init-only: ReadInputFile + LoadSourceConditions + writers, no GPU): every reference
input.  GPU part: whole runs through libkestrel_gpu, files equal to the Python host's."""
import glob
import os
import subprocess

import numpy as np
import pytest

from common import INPUTS
from kestrel_b200 import build as kbuild
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.run import Simulation, calculate_volume, write_solution_txt
from kestrel_b200.host.sources import load_source_conditions

CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(INPUTS, "case_*.txt")))


@pytest.fixture(scope="module")
def driver():
    if not os.path.exists(kbuild.LIB):
        pytest.skip("libkestrel_gpu.so not built")
    return kbuild.build_host()


def load_txt(path):
    rows = [ln for ln in open(path).read().splitlines() if ln.strip()]
    return np.array([[float(x) for x in ln.split(",")] for ln in rows])


def same(a, b, exact):
    assert a.shape == b.shape
    if exact:
        assert np.array_equal(a, b)
    else:  # numpy and libm sin/cos/tanh may differ in the last bit
        assert np.allclose(a, b, rtol=1e-12, atol=1e-13)


def libm_free(rs):
    return rs.topog_func in ("flat", "xslope", "yslope", "xyslope", "xparab", "xyparab")


@pytest.mark.parametrize("case", CASES)
def test_init_only_matches_python_host(driver, tmp_path, case):
    path = os.path.join(INPUTS, case)
    out = tmp_path / "cpp"
    r = subprocess.run([driver, path, "-o", str(out), "--init-only", "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rs = read_input_file(path)
    tiles = load_source_conditions(rs)
    snap = {tid: {"u": T.u} for tid, T in tiles.items()}
    ref = tmp_path / "py.txt"
    write_solution_txt(rs, str(ref), snap)
    a, b = load_txt(out / "000000.txt"), load_txt(ref)
    same(a, b, libm_free(rs))
    if libm_free(rs):  # the files themselves, byte for byte
        assert (out / "000000.txt").read_text() == ref.read_text()
    vol = calculate_volume(rs, snap)
    lines = (out / "Volume.txt").read_text().splitlines()
    assert lines[0].split(",")[0].strip() == "time" and len(lines) == 2
    got = [float(x) for x in lines[1].split(",")]
    assert np.allclose(got[1:], vol, rtol=1e-12, atol=1e-300)


def test_fatal_error_behaviour(driver, tmp_path):
    """A cap on an edge tile with bcs = halt: the reference's FatalErrorMessage (UpdateTiles.f90:63-65)."""
    bad = tmp_path / "bad.txt"
    bad.write_text("Domain:\nnXtiles = 3\nnYtiles = 3\nnXpertile = 4\nnYpertile = 4\nXtilesize = 4.0\n\n"
                   "Cap:\ncapX = -5.0\ncapY = -5.0\ncapRadius = 1.0\ncapHeight = 1.0\n\nSolver:\nT end = 1.0\n\n"
                   "Topog:\nType = Function\nTopog function = flat\n")
    r = subprocess.run([driver, str(bad), "-o", str(tmp_path / "o"), "--init-only"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "tried to add a tile outside the domain" in r.stderr
    r = subprocess.run([driver, str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open input file" in r.stderr


@pytest.mark.parametrize("func,params", [
    ("usgs", [0.3, 4.0]), ("flume", [28.0, 3.0, 6.0, 1.5, 0.4, 3.0]), ("channel power law", [-0.05, 4.0, 2.5]),
    ("channel trapezium", [-0.05, 5.0, 0.8]), ("xtrislope", [25.0, 10.0, 2.0, 4.0, -6.0, 5.0]), ("x2slopes", [0.5, 0.1, 12.0]),
    ("xbislope", [20.0, 5.0, 3.0]),
])
def test_every_topography_function_agrees_between_the_hosts(driver, tmp_path, func, params):
    """TopogFuncs.f90 in the C++ host (GetHeights) and in the Python host: the base elevation column of the initial
    output agrees to libm rounding for the functions the reference inputs do not exercise."""
    from kestrel_b200.host.inputfile import write_input_file
    from kestrel_b200.host.settings import Cap, RunSet
    rs = RunSet(nXtiles=5, nYtiles=5, nXpertile=12, nYpertile=10, Xtilesize=6.0, bcs="halt", topog_func=func, topog_params=params,
                tend=1.0, Nout=1)
    rs.caps = [Cap(x=0.0, y=0.0, radius=7.0, height=0.5, psi=0.0, shape="flat")]
    rs.finalize()
    inp = tmp_path / "in.txt"
    write_input_file(rs, str(inp))
    out = tmp_path / "cpp"
    r = subprocess.run([driver, str(inp), "-o", str(out), "--init-only", "--quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rs2 = read_input_file(str(inp))
    tiles = load_source_conditions(rs2)
    ref = tmp_path / "py.txt"
    write_solution_txt(rs2, str(ref), {tid: {"u": T.u} for tid, T in tiles.items()})
    a, b = load_txt(out / "000000.txt"), load_txt(ref)
    assert a.shape == b.shape and a.shape[0] > 0
    assert np.allclose(a, b, rtol=1e-9, atol=1e-11)      # 10 printed digits
    assert np.ptp(b[:, 11]) > 0.0                         # the base elevation column varies: the function was evaluated


@pytest.mark.gpu
@pytest.mark.parametrize("case,args", [
    ("case_1d_cap_constslope.txt", ["--tend", "20", "--nout", "2"]),
    ("case_cap_dilute_2d.txt", ["--tend", "1.0", "--nout", "2"]),
    ("case_flux_morpho.txt", ["--tend", "3.0", "--nout", "1"]),
    ("case_cap_morpho_2d.txt", ["--tend", "0.5", "--nout", "1", "--arithmetic", "1"]),
])
def test_run_matches_python_host(driver, gpu_lib, tmp_path, case, args):
    path = os.path.join(INPUTS, case)
    out = tmp_path / "cpp"
    r = subprocess.run([driver, path, "-o", str(out)] + args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rs = read_input_file(path)
    kv = dict(zip(args[::2], args[1::2]))
    rs.tend, rs.Nout, rs.arithmetic = float(kv["--tend"]), int(kv["--nout"]), int(kv.get("--arithmetic", 0))
    rs.finalize()
    py = tmp_path / "py"
    Simulation(rs, gpu_lib).run(out_dir=str(py))
    exact = libm_free(rs)
    for i in range(rs.Nout + 1):
        same(load_txt(out / f"{i:06d}.txt"), load_txt(py / f"{i:06d}.txt"), exact)
    va = [[float(x) for x in ln.split(",")] for ln in (out / "Volume.txt").read_text().splitlines()[1:]]
    vb = [[float(x) for x in ln.split(",")] for ln in (py / "Volume.txt").read_text().splitlines()[1:]]
    assert np.allclose(va, vb, rtol=1e-12, atol=1e-300)
