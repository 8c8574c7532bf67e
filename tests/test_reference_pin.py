"""Pins the oracle to the REAL reference -- on machines that can build it.

Activates itself when gfortran is on PATH and the reference sources are present (KESTREL_SRC or
/root/reference/src); skipped otherwise, as in the image this repository is developed in (no Fortran compiler
of any kind: DESIGN.md, oracle/ref_build/README.md).  oracle/ref_build/build_ref.sh compiles the reference's own
sources with GDAL / PROJ stubs into oracle/_ref/, a patched copy of Run dumps raw fp64 state after every output,
and the C++ oracle must reproduce those dumps: identical active-tile sets, fields to 1e-12 relative."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from common import INPUTS, ROOT, rel_linf

SRC = os.environ.get("KESTREL_SRC", "/root/reference/src")
RECIPE = os.path.join(ROOT, "oracle", "ref_build")
needs_ref = pytest.mark.skipif(shutil.which("gfortran") is None or not os.path.exists(os.path.join(SRC, "TimeStepper.f90")),
                               reason="no gfortran / no reference sources: the reference cannot be built here (parity unpinned)")


def test_recipe_is_complete():
    """Runs everywhere: the recipe's files exist, the stubs define exactly the symbols the reference binds, and the
    sed patch applies to the reference's TimeStepper.f90 when that file is available."""
    for f in ("README.md", "build_ref.sh", "gdal_proj_stubs.c", "raw_dump.f90", "patch_timestepper.sed", "read_raw.py"):
        assert os.path.exists(os.path.join(RECIPE, f)), f
    stubs = open(os.path.join(RECIPE, "gdal_proj_stubs.c")).read()
    bound = ["MallocDouble", "FreeDouble", "GeoTiffInfo", "GeoTiffArraySectionRead", "BuildDEMVRT_raster", "BuildDEMVRT_srtm",
             "proj_transformer__new", "proj_transformer__delete", "proj_transformer__wgs84_to_utm", "proj_transformer__utm_to_wgs84",
             "latlon_to_zone_number", "zone_number_to_central_longitude", "latlon_to_utm_epsg"]
    for s in bound:
        assert s + "(" in stubs, s
    r = subprocess.run(["gcc", "-Wall", "-Werror", "-c", os.path.join(RECIPE, "gdal_proj_stubs.c"), "-o", os.devnull], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ts = os.path.join(SRC, "TimeStepper.f90")
    if os.path.exists(ts):
        out = subprocess.run(["sed", "-f", os.path.join(RECIPE, "patch_timestepper.sed"), ts], capture_output=True, text=True).stdout
        assert out.count("call DumpRawState(RunParams, RunParams%CurrentOut, grid)") == 2
        assert out.count("use raw_dump_module, only: DumpRawState") == 1
        # every bind(C) name of the reference's two wrapper modules is covered by a stub
        import re
        names = set()
        for f in ("GeoTiffRead.f90", "utm.f90"):
            names |= set(re.findall(r'bind\(C, ?name="(\w+)"\)', open(os.path.join(SRC, f)).read()))
        assert names == set(bound), names ^ set(bound)


def test_raw_reader_round_trip(tmp_path):
    """read_raw.py against a file written in raw_dump.f90's layout."""
    sys.path.insert(0, RECIPE)
    from read_raw import read_raw
    nX, nY = 3, 2
    rng = np.random.default_rng(0)
    u = rng.random((nY, nX, 13)); b0 = rng.random((nY + 1, nX + 1)); bt = rng.random((nY + 1, nX + 1))
    mx = rng.random((5, 2, nY, nX)); tf = rng.random((nY, nX))
    with open(tmp_path / "raw_000001.bin", "wb") as fh:
        fh.write(np.array([1, nX, nY, 0], dtype="<i4").tobytes()); fh.write(np.array([2.5], dtype="<f8").tobytes())
        fh.write(np.array([7], dtype="<i4").tobytes())
        for a in (u, b0, bt, mx[0], mx[1], mx[2], mx[3], mx[4], tf):
            fh.write(a.astype("<f8").tobytes())
    t, tiles = read_raw(str(tmp_path / "raw_000001.bin"))
    assert t == 2.5 and list(tiles) == [7]
    assert np.array_equal(tiles[7]["u"], u) and np.array_equal(tiles[7]["maxima"], mx) and np.array_equal(tiles[7]["bt"], bt)


@needs_ref
@pytest.mark.parametrize("case,tend", [("lake_at_rest_hydro_2d", 5.0), ("flux_hydro_2d", 10.0), ("cap_morpho", 10.0), ("cap_dilute_2d", 8.0)])
def test_oracle_reproduces_the_reference(oracle_lib, tmp_path, case, tend):
    sys.path.insert(0, RECIPE)
    from read_raw import read_raw
    from common import run_input
    binary = subprocess.run([os.path.join(RECIPE, "build_ref.sh")], capture_output=True, text=True)
    assert binary.returncode == 0, binary.stderr
    exe = binary.stdout.strip().splitlines()[-1]
    ref_in = os.path.join(os.path.dirname(SRC), "tests", f"Input_{case}.txt")
    work = tmp_path / "run"
    work.mkdir()
    shutil.copy(ref_in, work / "Input.txt")
    r = subprocess.run([exe, "Input.txt"], cwd=work, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outdir = next(p for p in work.iterdir() if p.is_dir())
    dumps = sorted(p for p in outdir.iterdir() if p.name.startswith("raw_"))
    assert len(dumps) >= 2
    so = run_input(oracle_lib, os.path.join(INPUTS, f"case_{case}.txt"))
    assert len(so.snapshots) == len(dumps)
    for snap, dump in zip(so.snapshots, dumps):
        t, tiles = read_raw(str(dump))
        assert sorted(tiles) == sorted(snap), dump.name
        for d in range(13):
            A = np.concatenate([snap[k]["u"][..., d].ravel() for k in sorted(snap)])
            B = np.concatenate([tiles[k]["u"][..., d].ravel() for k in sorted(tiles)])
            assert np.max(np.abs(B)) < 1e-12 and np.max(np.abs(A)) < 1e-12 or rel_linf(A, B) <= 1e-12, (dump.name, d, rel_linf(A, B))
        for k in tiles:
            assert np.max(np.abs(snap[k]["bt"] - tiles[k]["bt"])) <= 1e-12 * max(1.0, np.max(np.abs(tiles[k]["bt"])))
