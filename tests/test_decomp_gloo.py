"""World-size-2 CPU test (gloo) of the host side of the multi-GPU path: block ownership, the
per-rank initial state, the neighbour map and the strip bookkeeping of the halo exchange,
exercised with torch.distributed send/recv on numpy planes laid out like the device planes.
(The CUDA/NCCL leg of the same exchange is tests/run_multigpu.py, needs >= 2 GPUs.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import ROOT
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state, decomposition, rank_block

H = 2  # halo width in cells (HydraulicRHS.f90:441-442: a limited slope needs the neighbour's second cell)


def neighbours(rank, px, py):
    rx, ry = rank % px, rank // px
    rk = lambda x, y: ((y + py) % py) * px + ((x + px) % px)
    return rk(rx - 1, ry), rk(rx + 1, ry), rk(rx, ry - 1), rk(rx, ry + 1)


def exchange(plane, rank, px, py):
    """The two-phase exchange of kgpu exchangeHalo on a padded numpy plane [NY+2H, NX+2H]."""
    west, east, south, north = neighbours(rank, px, py)
    NY, NX = plane.shape[0] - 2 * H, plane.shape[1] - 2 * H

    def swap(send_lo, send_hi, lo, hi):
        r_hi, r_lo = torch.empty_like(send_lo), torch.empty_like(send_hi)
        ops = [dist.P2POp(dist.isend, send_lo, lo), dist.P2POp(dist.isend, send_hi, hi),
               dist.P2POp(dist.irecv, r_hi, hi), dist.P2POp(dist.irecv, r_lo, lo)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        return r_lo, r_hi

    if px > 1:
        a = torch.from_numpy(np.ascontiguousarray(plane[H:H + NY, H:2 * H]))
        b = torch.from_numpy(np.ascontiguousarray(plane[H:H + NY, NX:NX + H]))
        r_lo, r_hi = swap(a, b, west, east)
        plane[H:H + NY, :H] = r_lo.numpy()
        plane[H:H + NY, NX + H:] = r_hi.numpy()
    else:
        plane[H:H + NY, :H] = plane[H:H + NY, NX:NX + H]
        plane[H:H + NY, NX + H:] = plane[H:H + NY, H:2 * H]
    if py > 1:
        a = torch.from_numpy(np.ascontiguousarray(plane[H:2 * H, :]))
        b = torch.from_numpy(np.ascontiguousarray(plane[NY:NY + H, :]))
        r_lo, r_hi = swap(a, b, south, north)
        plane[:H, :] = r_lo.numpy()
        plane[NY + H:, :] = r_hi.numpy()
    else:
        plane[:H, :] = plane[NY:NY + H, :]
        plane[NY + H:, :] = plane[H:2 * H, :]


def worker(rank, world, port, px, py, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = dambreak_runset(4, 8)
    blk = rank_block(rs, rank, px, py)
    q4, b0v = dambreak_state(rs, blk)
    NY, NX = q4.shape[1:]
    plane = np.zeros((NY + 2 * H, NX + 2 * H))
    plane[H:-H, H:-H] = q4[0]
    exchange(plane, rank, px, py)
    # the haloed block must equal the periodic window of the global field
    Q, _ = dambreak_state(rs)
    i0, j0 = blk[0] * rs.nXpertile, blk[1] * rs.nYpertile
    ii = (np.arange(-H, NX + H) + i0) % rs.NX
    jj = (np.arange(-H, NY + H) + j0) % rs.NY
    ok = np.array_equal(plane, Q[0][np.ix_(jj, ii)])
    # dt decision: one min-allreduce, identical on every rank
    local_min = torch.tensor([float(np.min(plane[H:-H, H:-H])) + rank])
    dist.all_reduce(local_min, op=dist.ReduceOp.MIN)
    results[rank] = (bool(ok), float(local_min.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("px,py", [(2, 1), (1, 2)])
def test_two_rank_halo_exchange_and_min_reduce(px, py):
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29600 + px * 10 + py
    mp.spawn(worker, args=(world, port, px, py, results), nprocs=world, join=True)
    assert all(results[r][0] for r in range(world)), dict(results)
    assert len({results[r][1] for r in range(world)}) == 1


def test_block_ownership_rule():
    rs = dambreak_runset(8, 16)
    for n in (1, 2, 4, 8):
        px, py = decomposition(n)
        assert px * py == n
        seen = set()
        for r in range(n):
            tx0, ty0, ntx, nty = rank_block(rs, r, px, py)
            for ty in range(ty0, ty0 + nty):
                for tx in range(tx0, tx0 + ntx):
                    assert (tx, ty) not in seen
                    seen.add((tx, ty))
        assert len(seen) == rs.nXtiles * rs.nYtiles


def test_dynamic_multigpu_command_lines_parse():
    """The command lines tests/test_gpu_multi.py builds for the dynamic-tile runs (reference inputs with overridden tile
    sizes, boundary conditions and end times) are parsed by the runner on CPU: run settings, decomposability of the tile
    grid over 2 x 1 / 2 x 2 / 4 x 1 ranks, and initial tiles that the host rasteriser actually produces."""
    import re
    import subprocess
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gpu_multi as tgm
    for case, extra in tgm.DYNAMIC_CASES:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "run_multigpu_dynamic.py"), "--case", case, "--dry-run"] + list(extra)
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
        m = re.search(r"DRY-RUN case=\S+ tiles=(\d+)x(\d+) of (\d+)x(\d+) bcs=(\w+) .* initial tiles=\[(.*)\]", r.stdout)
        assert m, r.stdout
        ntx, nty, nx, ny = (int(m.group(k)) for k in range(1, 5))
        assert ntx % 4 == 0 or (ntx % 2 == 0 and (nty == 1 or nty % 2 == 0)), (case, ntx, nty)
        assert nx >= 3 and (nty == 1 or ny >= 3)
        assert m.group(5) in ("halt", "dirichlet")
        assert len(m.group(6).split(",")) >= 2
