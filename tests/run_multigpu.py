"""Multi-GPU parity: a P-rank decomposed run must be BITWISE equal to the 1-GPU run.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
             --master-port 29511 tests/run_multigpu.py [--tiles T] [--per 64] [--steps 30]
The stencil is deterministic, halos are exact copies and the dt reduction is a true min, so
nothing may differ -- the analogue of the reference's tile-independence tests.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kestrel_b200 import capi  # noqa: E402
from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state, decomposition, rank_block, thin_dambreak_runset  # noqa: E402


def make_runset(args):
    if args.thin:
        args.morpho = True
        return thin_dambreak_runset(args.tiles, args.per)
    return dambreak_runset(args.tiles, args.per, morpho=args.morpho)


def attach(lib, st, rank, device):
    n = lib.comm_id_bytes()
    buf = torch.zeros(n, dtype=torch.uint8, device=device)
    if rank == 0:
        host = (capi.C.c_ubyte * n)()
        assert lib.comm_create_id(host) == 0
        buf.copy_(torch.tensor(list(host), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    idb = (capi.C.c_ubyte * n)(*buf.cpu().tolist())
    rc = lib.comm_attach(st.h, idb)
    assert rc == 0, lib.last_error(st.h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=4)
    ap.add_argument("--per", type=int, default=64)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--morpho", action="store_true", help="Strang-split run with the morphodynamic operator (bed included in the comparison)")
    ap.add_argument("--thin", action="store_true", help="thin-layer morphodynamic dam-break: RedistributeGrid runs every step (across ranks)")
    ap.add_argument("--arithmetic", type=int, default=0)
    args = ap.parse_args()
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    lib = capi.load_gpu()
    px, py = decomposition(world)
    rs = make_runset(args)
    rs.device = lrank
    rs.arithmetic = args.arithmetic
    rs.comm_rank, rs.comm_size, rs.comm_px, rs.comm_py = rank, world, px, py
    blk = rank_block(rs, rank, px, py)
    q4, b0v = dambreak_state(rs, blk)
    p, keep = rs.to_c()
    st = capi.Stepper(lib, p, keep)
    attach(lib, st, rank, dev)
    st.upload_domain(q4, b0v)
    info = st.integrate_to(1e30, args.steps)
    q, bt = st.download_domain(want_bt=True)
    # gather blocks on rank 0
    t = torch.from_numpy(q).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, parts, 0)
    tb = torch.from_numpy(np.ascontiguousarray(bt)).to(dev)
    bparts = [torch.empty_like(tb) for _ in range(world)] if rank == 0 else None
    dist.gather(tb, bparts, 0)
    ok = True
    if rank == 0:
        rs1 = make_runset(args)
        rs1.device = lrank
        rs1.arithmetic = args.arithmetic
        Q4, B0 = dambreak_state(rs1)
        p1, k1 = rs1.to_c()
        s1 = capi.Stepper(lib, p1, k1)
        s1.upload_domain(Q4, B0)
        i1 = s1.integrate_to(1e30, args.steps)
        ref, refbt = s1.download_domain(want_bt=True)
        got = np.empty_like(ref)
        gotbt = np.empty_like(refbt)
        for r in range(world):
            tx0, ty0, ntx, nty = rank_block(rs1, r, px, py)
            got[:, ty0 * args.per:(ty0 + nty) * args.per, tx0 * args.per:(tx0 + ntx) * args.per] = parts[r].cpu().numpy()
            # seam vertices are held by both neighbours (and must agree): later blocks overwrite with equal values
            gotbt[ty0 * args.per:(ty0 + nty) * args.per + 1, tx0 * args.per:(tx0 + ntx) * args.per + 1] = bparts[r].cpu().numpy()
        same_t = (i1.t == info.t and i1.nsteps == info.nsteps and i1.nrefines == info.nrefines)
        exact = [bool(np.array_equal(got[d], ref[d])) for d in range(4)]
        err = [float(np.max(np.abs(got[d] - ref[d]))) for d in range(4)]
        bed_exact = bool(np.array_equal(gotbt, refbt))
        bed_moved = float(np.max(np.abs(refbt)))
        ok = same_t and all(exact) and bed_exact and (bed_moved > 0.0 or not args.morpho)
        print(f"MULTIGPU world={world} decomposition={px}x{py} grid={rs1.NX}x{rs1.NY} morpho={args.morpho} steps={info.nsteps} "
              f"refines={info.nrefines} t={info.t!r} same_t={same_t} bitwise={exact} maxabs={err} bed_bitwise={bed_exact} "
              f"max|bt|={bed_moved:.3e} -> {'PASS' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    st.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
