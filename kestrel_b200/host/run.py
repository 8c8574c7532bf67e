"""Host-side Run loop around the C-ABI (TimeStepper.f90:73-113).

The Fortran host's Run does: output, then for each output interval
IntegrateTo(t_k) followed by OutputSolutionData / CalculateVolume.  Here
IntegrateTo is the library call ``integrate_to``; the writers reproduce
Volume.txt (Output.f90:617-735) and the per-cell txt layout the reference's
tests index (Output.f90:799-834; tests/testlib.jl:463-474).
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional

import numpy as np

from .. import capi
from .settings import RunSet
from .sources import W, HU, HV, HPSI, HN, U, V, PSI, RHO, B0, BT, BX, BY, load_source_conditions
from .topog import make_heights_callback, tile_coords


def kahan_total(values: np.ndarray) -> float:
    s = 0.0
    c = 0.0
    for x in values:
        y = x - c
        t = s + y
        c = (t - s) - y
        s = t
    return s


def calculate_volume(rs: RunSet, tiles: Dict[int, dict]):
    """CalculateVolume (Output.f90:617-735) over downloaded active tiles.
    Returns (vol, bed, mass, bed*rhob, solids*rhos, bed*rhos*(1-p))."""
    dv_all, mass_all, bed_all, sol_all = [np.zeros(0)], [np.zeros(0)], [np.zeros(0)], [np.zeros(0)]
    for tid in sorted(tiles):
        u = tiles[tid]["u"]
        gam = np.sqrt(1.0 + u[..., BX] * u[..., BX] + u[..., BY] * u[..., BY]) if rs.geometric_factors else np.ones_like(u[..., BX])
        dv_all.append(((u[..., W] - u[..., B0] - u[..., BT]) * gam * gam).ravel())
        mass_all.append((u[..., RHO] * u[..., HN] * gam).ravel())
        bed_all.append(u[..., BT].ravel())
        sol_all.append((u[..., HPSI] * gam).ravel())
    area = rs.deltaX if rs.isOneD else rs.deltaX * rs.deltaY
    vol = kahan_total(np.concatenate(dv_all)) * area
    mass = kahan_total(np.concatenate(mass_all)) * area
    bed = kahan_total(np.concatenate(bed_all)) * area
    solids = kahan_total(np.concatenate(sol_all)) * area
    rhob = rs.rhow * rs.BedPorosity + rs.rhos * (1.0 - rs.BedPorosity)
    return (vol, bed, mass, bed * rhob, solids * rs.rhos, bed * rs.rhos * (1.0 - rs.BedPorosity))


def write_solution_txt(rs: RunSet, path: str, tiles: Dict[int, dict]):
    """OutputSolutionData_txt column layout (Output.f90:799-834):
    1-D: tile, x, Hn, w, u, speed, density, base_elev, Hnpsi, psi, rhoHnu, bt ... (Hn=3,u=5,Hnpsi=9,bt=12)
    2-D: tile, x, y, lat, lon, Hn, w, u, v, speed, density, b0, bt_c?, Hnpsi ... (Hn=6,u=8,Hnpsi=14,bt=17)
    Only the columns the reference tests index are guaranteed to line up."""
    with open(path, "w") as fh:
        for tid in sorted(tiles):
            u = tiles[tid]["u"]
            x, y, _, _ = tile_coords(rs, tid)
            nY, nX = u.shape[:2]
            for j in range(nY):
                for i in range(nX):
                    q = u[j, i]
                    spd = np.sqrt(q[U] ** 2 + q[V] ** 2)
                    if rs.isOneD:
                        cols = [tid, x[i], q[HN], q[W], q[U], spd, q[RHO], q[B0] + q[BT], q[HPSI], q[PSI], q[HU], q[BT], q[BX], q[B0]]
                    else:
                        cols = [tid, x[i], y[j], 0.0, 0.0, q[HN], q[W], q[U], q[V], spd, q[RHO], q[B0], q[B0] + q[BT],
                                q[HPSI], q[PSI], q[HU], q[BT], q[HV], q[BX], q[BY]]
                    fh.write(", ".join(f"{c:18.10E}" if isinstance(c, float) or isinstance(c, np.floating) else f"{c:8d}" for c in cols) + "\n")
                fh.write("\n")


class Simulation:
    """LoadSourceConditions + Run against a library exporting the ABI."""

    def __init__(self, rs: RunSet, lib: capi.Library, device_topography: bool = False,
                 after_create: Optional[Callable] = None):
        """after_create(stepper): called between kgpu_create and the first upload -- a decomposed run attaches its
        communicator there (kgpu_comm_attach).  Every rank of a decomposed run builds the same initial tiles and
        uploads all of them (kgpu_upload_tile is collective when tiles are dynamic); downloads cover its own block."""
        self.rs = rs
        self.lib = lib
        self.ic_tiles = load_source_conditions(rs)  # also fills NumCellsInSrc
        p, keep = rs.to_c(make_heights_callback(rs))
        self.stepper = capi.Stepper(lib, p, keep)
        if after_create:
            after_create(self.stepper)
        if device_topography:  # tiles activated during the run get their heights from a kernel, not from the callback
            self.stepper.set_topography_function(rs.topog_func, rs.topog_params)
        for tid in sorted(self.ic_tiles):
            T = self.ic_tiles[tid]
            self.stepper.upload_tile(tid, T.u, b0v=np.ascontiguousarray(T.b0v), maxima=T.maxima, tfirst=T.tfirst,
                                     contains_source=T.contains_source)
        self.volume_rows: List[tuple] = []
        self.snapshots: List[Dict[int, dict]] = []
        self.infos: List[capi.KgpuStepInfo] = []

    def download_active(self) -> Dict[int, dict]:
        return {int(t): self.stepper.download_tile(int(t)) for t in self.stepper.active_tiles()}

    def initial_tiles(self) -> Dict[int, dict]:
        own = set(int(t) for t in self.stepper.active_tiles()) if self.rs.comm_size > 1 else None
        return {tid: {"u": T.u.copy(), "b0": T.b0v, "bt": np.zeros_like(T.b0v), "maxima": T.maxima, "tfirst": T.tfirst}
                for tid, T in self.ic_tiles.items() if own is None or tid in own}

    def run(self, out_dir: Optional[str] = None, keep_snapshots: bool = True,
            on_output: Optional[Callable] = None):
        rs = self.rs
        tiles0 = self.initial_tiles()
        self.volume_rows.append((rs.tstart,) + calculate_volume(rs, tiles0))
        if keep_snapshots:
            self.snapshots.append(tiles0)
        if out_dir:
            os.makedirs(out_dir, exist_ok=True)
            write_solution_txt(rs, os.path.join(out_dir, "000000.txt"), tiles0)
        for i in range(1, rs.Nout + 1):
            tk = rs.tstart + i * rs.DeltaT
            info = self.stepper.integrate_to(tk)
            self.infos.append(info)
            tiles = self.download_active()
            self.volume_rows.append((tk,) + calculate_volume(rs, tiles))
            if keep_snapshots:
                self.snapshots.append(tiles)
            if out_dir:
                write_solution_txt(rs, os.path.join(out_dir, f"{i:06d}.txt"), tiles)
            if on_output:
                on_output(i, tk, tiles)
        if out_dir:
            with open(os.path.join(out_dir, "Volume.txt"), "w") as fh:
                fh.write("        time,                   volume,         total_bed_volume,               total_mass,"
                         "                 bed_mass,        total_solids_mass,          bed_solids_mass\n")
                for row in self.volume_rows:
                    fh.write(f"{row[0]:12.2f}" + "".join(f", {v:24.15E}" for v in row[1:]) + "\n")
        return self
