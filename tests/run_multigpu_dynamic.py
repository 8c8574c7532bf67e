"""Multi-GPU parity with DYNAMIC tiles: a P-rank decomposed run of a reference input (non-periodic domain, tiles
activated by the flow, ghost tiles, flux sources) must be BITWISE equal to the 1-GPU run -- same active and ghost
sets at every output, same step / refinement / tile counts, same 13 fields, maxima and heights per tile.

Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 \
             --master-port 29561 tests/run_multigpu_dynamic.py --case case_flux_hydro_2d.txt [--set "T end=4.0"] ...
Every rank reads the input file, builds the same initial tiles on the host and uploads all of them (kgpu_upload_tile is
collective in this mode: the replicated tile table, kgpu_tile_table.hpp, must see the same mutations everywhere);
each rank downloads the tiles of its own block and rank 0 compares their union with its own single-device run.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kestrel_b200 import capi  # noqa: E402
from kestrel_b200.host.inputfile import read_input_file  # noqa: E402
from kestrel_b200.host.run import Simulation  # noqa: E402
from kestrel_b200.host.synthetic import decomposition  # noqa: E402
from run_multigpu import attach  # noqa: E402

INPUTS = os.path.join(ROOT, "tests", "inputs")


def make_runset(args):
    rs = read_input_file(os.path.join(INPUTS, args.case))
    for kv in args.set or []:
        k, v = kv.split("=", 1)
        cur = getattr(rs, k)
        if v == "None":
            setattr(rs, k, None)
        elif isinstance(cur, bool):
            setattr(rs, k, v.lower() in ("1", "true", "on"))
        else:
            setattr(rs, k, float(v) if cur is None else type(cur)(v))
    rs.arithmetic = args.arithmetic
    return rs


def pack(sim):
    """What a rank contributes: per output, its own tiles; plus the step counters and its ghost tiles at the end."""
    snaps = [{tid: {k: np.ascontiguousarray(v) for k, v in d.items()} for tid, d in s.items()} for s in sim.snapshots]
    infos = [(i.t, i.nsteps, i.nrefines, i.ntiles_added) for i in sim.infos]
    return snaps, infos, sorted(int(t) for t in sim.stepper.ghost_tiles())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="case_flux_hydro_2d.txt")
    ap.add_argument("--set", action="append", help="RunSet attribute override, e.g. --set Nout=2 --set DeltaT=1.5")
    ap.add_argument("--arithmetic", type=int, default=0)
    ap.add_argument("--px", type=int, default=0)
    ap.add_argument("--expect-halt", action="store_true",
                    help="the flow reaches the domain edge under Boundary Conditions = halt: every rank and the single device must stop "
                         "with KGPU_ERR_HALT_BC (UpdateTiles.f90:63-65) after the same number of steps")
    ap.add_argument("--dry-run", action="store_true", help="parse the arguments, build the run settings and the initial tiles on the host, exit")
    args = ap.parse_args()
    if args.dry_run:   # CPU check of the command lines tests/test_gpu_multi.py builds (tests/test_decomp_gloo.py)
        from kestrel_b200.host.sources import load_source_conditions
        rs = make_runset(args)
        rs.finalize()
        tiles = load_source_conditions(rs)
        print(f"DRY-RUN case={args.case} tiles={rs.nXtiles}x{rs.nYtiles} of {rs.nXpertile}x{rs.nYpertile} bcs={rs.bcs} tend={rs.tend} "
              f"Nout={rs.Nout} morpho={rs.MorphodynamicsOn} initial tiles={sorted(tiles)}")
        return
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    lib = capi.load_gpu()
    px, py = decomposition(world)
    if args.px:
        px, py = args.px, world // args.px
    rs = make_runset(args)
    if rs.isOneD:
        px, py = world, 1
    rs.device = lrank
    rs.comm_rank, rs.comm_size, rs.comm_px, rs.comm_py = rank, world, px, py
    rs.finalize()
    sim = Simulation(rs, lib, after_create=lambda st: attach(lib, st, rank, dev))
    if args.expect_halt:
        def until_error(s):
            try:
                s.run()
            except capi.KestrelError as e:
                return (e.code, [(i.t, i.nsteps, i.nrefines, i.ntiles_added) for i in s.infos], sorted(int(t) for t in s.stepper.active_tiles()))
            return (0, [], [])
        mine = until_error(sim)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, 0)
        ok = True
        if rank == 0:
            rs1 = make_runset(args)
            rs1.device = lrank
            rs1.finalize()
            ref = until_error(Simulation(rs1, lib))
            union = sorted(t for r in range(world) for t in parts[r][2])
            ok = ref[0] == 3 and all(p[0] == 3 and p[1] == ref[1] for p in parts) and union == ref[2]
            print(f"MULTIGPU-DYNAMIC world={world} decomposition={px}x{py} case={args.case} expect-halt: codes={[p[0] for p in parts]} "
                  f"reference={ref[0]} completed outputs={len(ref[1])} active tiles at the stop={len(ref[2])} -> {'PASS' if ok else 'FAIL'}", flush=True)
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        sim.stepper.close()
        dist.destroy_process_group()
        sys.exit(0 if int(flag.item()) == 1 else 1)
    sim.run()
    mine = pack(sim)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(mine, parts, 0)
    ok = True
    if rank == 0:
        rs1 = make_runset(args)
        rs1.device = lrank
        rs1.finalize()
        ref = Simulation(rs1, lib)
        ref.run()
        rsn, rinfos, rghost = pack(ref)
        msgs = []
        for r in range(world):
            if parts[r][1] != rinfos:
                ok = False
                msgs.append(f"rank {r} counters {parts[r][1]} != {rinfos}")
        ghosts = sorted(t for r in range(world) for t in parts[r][2])
        if ghosts != rghost:
            ok = False
            msgs.append(f"ghost tiles differ: {len(ghosts)} vs {len(rghost)}")
        ntiles = []
        for k, rsnap in enumerate(rsn):
            got = {}
            for r in range(world):
                for tid, d in parts[r][0][k].items():
                    if tid in got:
                        ok = False
                        msgs.append(f"output {k}: tile {tid} reported by two ranks")
                    got[tid] = d
            if sorted(got) != sorted(rsnap):
                ok = False
                msgs.append(f"output {k}: active tiles differ ({len(got)} vs {len(rsnap)})")
                continue
            ntiles.append(len(got))
            for tid in sorted(got):
                for name in ("u", "b0", "bt", "maxima", "tfirst"):
                    a, b = got[tid][name], rsnap[tid][name]
                    if not np.array_equal(a, b):
                        ok = False
                        if len(msgs) < 12:
                            msgs.append(f"output {k} tile {tid} {name}: max|diff| = {float(np.max(np.abs(a - b))):.3e}")
        grew = len(ntiles) > 1 and ntiles[-1] > ntiles[0]
        ok = ok and grew
        print(f"MULTIGPU-DYNAMIC world={world} decomposition={px}x{py} case={args.case} arithmetic={args.arithmetic} "
              f"counters(t, steps, refines, tiles added)={rinfos[-1]} active tiles per output={ntiles} ghosts={len(rghost)} "
              f"-> {'PASS' if ok else 'FAIL'}", flush=True)
        for m in msgs:
            print("   ", m, flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    sim.stepper.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
