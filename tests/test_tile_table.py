"""The replicated tile table (kestrel_b200/csrc/kgpu_tile_table.hpp) -- groundwork for dynamic tile activation
across ranks -- against the oracle, on CPU.  The oracle is driven one step at a time on the reference's
dynamic-tile inputs; before every step the four flag bits per active tile are computed from its downloaded
state exactly as Near{N,S,E,W}Boundary do (TimeStepper.f90:948-1150), fed to the table's replay of
CheckIfNearBoundaries, and the table's active set (ascending) and ghost set (creation order) must equal the
oracle's after the step -- including quirk Q3 (trip count fixed at loop entry while the list grows).
World size 2 over gloo: each rank computes the flags of its own half of the tile grid only, the halves are
all-reduced, and both ranks must arrive at the oracle's sets."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import INPUTS
from kestrel_b200 import build as kbuild
from kestrel_b200 import capi
from kestrel_b200.host.inputfile import read_input_file
from kestrel_b200.host.run import Simulation

I32 = C.POINTER(C.c_int32)
HN = 4


class Table:
    def __init__(self, rs):
        self.dll = C.CDLL(kbuild.build())
        d = self.dll
        d.kgpu_debug_tiletable_new.restype = C.c_void_p
        d.kgpu_debug_tiletable_new.argtypes = [C.c_int32] * 5
        d.kgpu_debug_tiletable_free.argtypes = [C.c_void_p]
        d.kgpu_debug_tiletable_add.argtypes = [C.c_void_p, C.c_int32]
        d.kgpu_debug_tiletable_replay.argtypes = [C.c_void_p, I32, C.c_int32, C.c_int32, C.c_int32]
        d.kgpu_debug_tiletable_lists.argtypes = [C.c_void_p, I32, I32, I32, I32, C.POINTER(C.c_int64), I32]
        self.rs = rs
        self.n = rs.nXtiles * (1 if rs.isOneD else rs.nYtiles)
        self.t = d.kgpu_debug_tiletable_new(rs.nXtiles, 1 if rs.isOneD else rs.nYtiles, int(rs.bcs == "periodic"), int(rs.isOneD),
                                            int(rs.bcs == "halt"))
        assert self.t

    def add(self, tid):
        assert self.dll.kgpu_debug_tiletable_add(self.t, int(tid)) == 0

    def replay(self, flags):
        f = np.ascontiguousarray(flags, dtype=np.int32)
        return self.dll.kgpu_debug_tiletable_replay(self.t, f.ctypes.data_as(I32), self.rs.nXpertile, self.rs.nYpertile, self.rs.TileBuffer)

    def lists(self):
        na, ng, nops, nadd = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        a, g = np.zeros(self.n, np.int32), np.zeros(self.n, np.int32)
        assert self.dll.kgpu_debug_tiletable_lists(self.t, C.byref(na), a.ctypes.data_as(I32), C.byref(ng), g.ctypes.data_as(I32),
                                                   C.byref(nadd), C.byref(nops)) == 0
        return a[:na.value].tolist(), g[:ng.value].tolist(), int(nadd.value), int(nops.value)

    def close(self):
        self.dll.kgpu_debug_tiletable_free(self.t)


def tile_flags(rs, u13):
    """Near{N,S,E,W}Boundary of one tile: bit 0 N, 1 S, 2 E, 3 W (kgpu_tiles.cuh tile_flags_kernel)."""
    wet = u13[..., HN] > rs.heightThreshold          # [nY, nX]
    buf, nY, nX = rs.TileBuffer, rs.nYpertile, rs.nXpertile
    f = 0
    if not rs.isOneD:
        f |= 1 if wet[max(nY - buf, 0):, :].any() else 0
        f |= 2 if wet[:buf, :].any() else 0
    f |= 4 if wet[:, max(nX - buf, 0):].any() else 0
    f |= 8 if wet[:, :buf].any() else 0
    return f


def oracle_sets(st):
    n, ng = C.c_int32(), C.c_int32()
    cap = st.nXt * st.nYt
    g = np.zeros(cap, np.int32)
    st.lib.ghost_tiles(st.h, C.byref(ng), g.ctypes.data_as(I32))
    return [int(x) for x in st.active_tiles()], sorted(int(x) for x in g[:ng.value])


def drive(case, kw, steps, oracle_lib, owner=None, reduce_flags=None):
    rs = read_input_file(os.path.join(INPUTS, case))
    for k, v in kw.items():
        setattr(rs, k, v)
    rs.finalize()
    sim = Simulation(rs, oracle_lib)
    st = sim.stepper
    T = Table(rs)
    for tid in sorted(sim.ic_tiles):
        T.add(tid)
    act, gh = oracle_sets(st)
    ta, tg, _, _ = T.lists()
    assert ta == act and sorted(tg) == gh, "after LoadSourceConditions"
    grew = 0
    for step in range(steps):
        flags = np.zeros(T.n, np.int32)
        for tid in st.active_tiles():
            if owner is None or owner(int(tid)):
                flags[int(tid) - 1] = tile_flags(rs, st.download_tile(int(tid))["u"])
        if reduce_flags is not None:
            flags = reduce_flags(flags)
        assert T.replay(flags) == 0
        st.integrate_to(1e30, 1)
        act, gh = oracle_sets(st)
        ta, tg, nadd, nops = T.lists()
        assert ta == act, f"active set after step {step + 1}"
        assert sorted(tg) == gh, f"ghost set after step {step + 1}"
        grew = nadd
    T.close()
    return grew


CASES = [
    ("case_tile_indep_dynamic_20m.txt", dict(), 160),
    ("case_cap_dilute.txt", dict(), 700),
    ("case_cap_morpho.txt", dict(), 250),
    ("case_flux_hydro_2d.txt", dict(), 120),
]


@pytest.mark.parametrize("case,kw,steps", CASES)
def test_replay_matches_the_oracle(oracle_lib, case, kw, steps):
    drive(case, kw, steps, oracle_lib)


def test_some_case_really_adds_tiles(oracle_lib):
    assert drive("case_tile_indep_dynamic_20m.txt", dict(), 160, oracle_lib) > 0


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = capi.Library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "libkestrel_oracle.so"), "kor_")
        rs0 = read_input_file(os.path.join(INPUTS, "case_tile_indep_dynamic_20m.txt")).finalize()
        half = rs0.nXtiles // 2

        def owner(tid):   # 2 x 1 decomposition of the tile grid: rank 0 owns the west half
            return ((tid - 1) % rs0.nXtiles < half) == (rank == 0)

        def reduce_flags(f):   # disjoint contributions: a sum is the all-gather
            t = torch.from_numpy(f.astype(np.int64))
            dist.all_reduce(t)
            return t.numpy().astype(np.int32)

        grew = drive("case_tile_indep_dynamic_20m.txt", dict(), 160, lib, owner, reduce_flags)
        q.put((rank, grew > 0))
    finally:
        dist.destroy_process_group()


def test_two_ranks_replay_the_same_table_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
