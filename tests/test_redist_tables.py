"""Host bookkeeping of RedistributeGrid across ranks (kestrel_b200/csrc/kgpu_redist_tables.hpp), exercised on
CPU through the kgpu_debug_redist_tables probe: walk order of the reference on GLOBAL indices
(Redistribute.f90:69-101), canonical patch slots with periodic wrap, and -- world size 2 over gloo -- that
every rank builds the same tables from the gathered lists (which is what makes the replicated walk identical)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kestrel_b200 import build as kbuild

I32 = C.POINTER(C.c_int32)
F64 = C.POINTER(C.c_double)


def _probe():
    dll = C.CDLL(kbuild.build())
    fn = dll.kgpu_debug_redist_tables
    fn.restype = C.c_int
    fn.argtypes = [I32, I32, F64, I32, I32, I32, I32, I32, I32, I32]
    return fn


def tables(geom, counts, excess, li, lj):
    fn = _probe()
    g = np.asarray(geom, dtype=np.int32)
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    excess = np.ascontiguousarray(excess, dtype=np.float64)
    li, lj = np.ascontiguousarray(li, dtype=np.int32), np.ascontiguousarray(lj, dtype=np.int32)
    n = int(counts.sum())
    nout = C.c_int32()
    patch, vslot, cslot = np.zeros(n, np.int32), np.zeros(n * 16, np.int32), np.zeros(n * 9, np.int32)
    uniq = np.zeros(2, np.int32)
    p = lambda a, t: a.ctypes.data_as(t)
    rc = fn(p(g, I32), p(counts, I32), p(excess, F64), p(li, I32), p(lj, I32), C.byref(nout), p(patch, I32), p(vslot, I32), p(cslot, I32), p(uniq, I32))
    assert rc == 0 and nout.value == n
    return patch, vslot.reshape(n, 16), cslot.reshape(n, 9), uniq


def random_lists(rng, R, M, NX, NY, ties=True):
    counts = rng.integers(0, M + 1, size=R)
    excess = rng.random(R * M)
    if ties:  # equal excesses force the scan-order tie break
        excess = np.round(excess, 1)
    li, lj = np.zeros(R * M, np.int32), np.zeros(R * M, np.int32)
    for r in range(R):  # distinct cells per rank
        cells = rng.choice(NX * NY, size=M, replace=False)
        li[r * M:(r + 1) * M], lj[r * M:(r + 1) * M] = cells % NX, cells // NX
    return counts, excess, li, lj


@pytest.mark.parametrize("px,py", [(1, 1), (2, 1), (2, 2), (4, 2)])
def test_walk_order_and_canonical_slots(px, py):
    rng = np.random.default_rng(7 + px * 10 + py)
    nX = nY = 8
    gnXt, gnYt = 4 * px, 2 * py
    NX, NY = nX * gnXt // px, nY * gnYt // py
    NXg, NYg = nX * gnXt, nY * gnYt
    R, M = px * py, 40
    counts, excess, li, lj = random_lists(rng, R, M, NX, NY)
    patch, vslot, cslot, uniq = tables([R, M, px, NX, NY, nX, nY, gnXt, gnYt, 0], counts, excess, li, lj)
    # global cells of the gathered entries
    r = patch // M
    gi, gj = li[patch] + (r % px) * NX, lj[patch] + (r // px) * NY
    assert sorted(patch.tolist()) == sorted(s for rr in range(R) for s in range(rr * M, rr * M + counts[rr]))
    key = list(zip(excess[patch], (gi // nX) + (gj // nY) * gnXt, gj, gi))
    assert key == sorted(key), "ascending excess, ties by global tile, then j, then i"
    # every slot points at a copy of the right vertex / cell, and one key has one slot
    for slots, width, n_per, first in ((vslot, 4, 16, 0), (cslot, 3, 9, 48)):
        seen = {}
        for e in range(len(patch)):
            for q in range(n_per):
                k = ((gi[e] - 1 + q % width) % NXg, (gj[e] - 1 + q // width) % NYg)
                s = int(slots[e, q])
                p2, off = divmod(s, 93)
                q2 = off - first
                assert 0 <= q2 < n_per
                e2 = int(np.nonzero(patch == p2)[0][0])
                assert ((gi[e2] - 1 + q2 % width) % NXg, (gj[e2] - 1 + q2 // width) % NYg) == k
                assert e2 <= e, "canonical copy = first patch of the walk that holds it"
                assert seen.setdefault(k, s) == s
        assert len(seen) == uniq[0 if n_per == 16 else 1]


def test_one_dimensional_patches():
    rng = np.random.default_rng(3)
    R, M, NX = 2, 10, 64
    counts = np.array([7, 4])
    excess = rng.random(R * M)
    li = np.concatenate([rng.choice(NX, M, replace=False), rng.choice(NX, M, replace=False)]).astype(np.int32)
    lj = np.zeros(R * M, np.int32)
    patch, vslot, cslot, uniq = tables([R, M, 2, NX, 1, 16, 1, 8, 1, 1], counts, excess, li, lj)
    gi = li[patch] + (patch // M) * NX
    for e in range(len(patch)):
        for q in range(16):   # every row of a 1-D patch is row 0
            p2, off = divmod(int(vslot[e, q]), 93)
            e2 = int(np.nonzero(patch == p2)[0][0])
            assert (gi[e2] - 1 + off % 4) % 128 == (gi[e] - 1 + q % 4) % 128


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        px, py, nX, nY, gnXt, gnYt, M = 2, 1, 8, 8, 8, 4, 30
        NX, NY = nX * gnXt // px, nY * gnYt // py
        rng = np.random.default_rng(100 + rank)          # every rank knows only its own list ...
        n_loc = int(rng.integers(5, M))
        cells = rng.choice(NX * NY, size=M, replace=False)
        mine = torch.zeros(M, 3, dtype=torch.float64)
        mine[:, 0] = torch.from_numpy(np.round(rng.random(M), 1))
        mine[:, 1] = torch.from_numpy((cells % NX).astype(np.float64))
        mine[:, 2] = torch.from_numpy((cells // NX).astype(np.float64))
        cnt = torch.tensor([n_loc], dtype=torch.int32)
        allc = [torch.zeros(1, dtype=torch.int32) for _ in range(world)]
        alle = [torch.zeros(M, 3, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allc, cnt)                        # ... the gather of redistributeAcrossRanks
        dist.all_gather(alle, mine)
        counts = torch.cat(allc).numpy()
        E = torch.cat(alle).numpy()
        patch, vslot, cslot, uniq = tables([world, M, px, NX, NY, nX, nY, gnXt, gnYt, 0], counts, E[:, 0], E[:, 1].astype(np.int32), E[:, 2].astype(np.int32))
        digest = torch.tensor([int(patch.sum()), int(vslot.astype(np.int64).sum()), int(cslot.astype(np.int64).sum()), int(uniq[0]), int(uniq[1]),
                               len(patch)], dtype=torch.int64)
        both = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(both, digest)
        ok = all(torch.equal(both[0], b) for b in both) and len(patch) == int(counts.sum())
        # an entry on the seam column of rank 1 shares vertices with rank 0's last column: one slot for both
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_every_rank_builds_identical_tables_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
