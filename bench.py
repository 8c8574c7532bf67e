#!/usr/bin/env python
"""bench.py -- cell-updates/s of the fused explicit time step on the synthetic dam-break.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl ours|reference]

One "step" = one complete time step of the do-while loop of IntegrateTo
(TimeStepper.f90:151-275): 4 fused RHS+stage kernels, the stage-1 update, maxima tracking and
dt control.  Workload (BASELINE.json configs[4]): periodic S x S dam-break on
xySinSlope topography, Chezy drag, erosion off, every tile active; S = 16384 on one B200 (each
fp64 field is 2.1 GB, far beyond the 126 MB L2, so no L2 flush is needed between steps).

Prints ONE JSON line.  value = whole-job cell-updates/s with the state resident in HBM, timed
with CUDA events on the library's stream.  e2e = the same metric through the C-ABI with HOST
buffers: kgpu_upload_domain (pinned host -> device) + K steps + kgpu_download_domain inside the
timed region -- one output interval of the reference's Run loop (TimeStepper.f90:99-111).
roofline = algorithmic bytes of the fused stage kernel (104 B/cell/launch) over its mean launch
duration measured live with CUDA events.  cpu_baseline = the CPU oracle (a port; the Fortran
reference cannot be built here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ALG_BYTES_PER_CELL_STAGE = 104   # read q(4)+q0(4)+b0(1), write q(4) fp64
ALG_BYTES_PER_CELL_UPDATE = 416  # 4 stages
CPU_SAMPLE_SIZE = 1024


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int = 0):
        self.samples, self.reasons, self.proc = [], set(), None
        self.index = index
        self.max_mhz = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            try:
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def oracle_library():
    from kestrel_b200 import capi
    path = os.path.join(ROOT, "oracle", "libkestrel_oracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)  # no-op when up to date
    return capi.Library(path, "kor_")


def cpu_baseline(steps: int, warmup: int, size: int = CPU_SAMPLE_SIZE, threads: int = 0, morpho: bool = False):
    """The oracle on the host cores, bounded sample of the same workload."""
    from kestrel_b200 import capi
    from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state
    lib = oracle_library()
    cores = threads or os.cpu_count() or 1
    rs = dambreak_runset(size // 128, 128, morpho=morpho)
    q4, b0v = dambreak_state(rs)
    p, keep = rs.to_c()
    st = capi.Stepper(lib, p, keep)
    lib.set_threads(st.h, cores)
    st.upload_domain(q4, b0v)
    st.integrate_to(1e30, warmup)
    t0 = time.perf_counter()
    st.integrate_to(1e30, steps)
    dt = time.perf_counter() - t0
    st.close()
    cells = rs.NX * rs.NY
    return {"value": cells * steps / dt, "unit": "cell-updates/s", "cores": cores, "kind": "port",
            "sample": f"{size}x{size} cells of the same dam-break, {warmup} warm-up + {steps} timed steps, OpenMP over rows "
                      f"({cores} threads); the Fortran reference is serial and cannot be built in this image",
            "ms_per_step": 1e3 * dt / steps}


def _new_id(lib):
    from kestrel_b200 import capi
    n = lib.comm_id_bytes()
    hb = (capi.C.c_ubyte * n)()
    assert lib.comm_create_id(hb) == 0
    return hb


def decomposed_bitwise_check(lib, dist, torch, rank, world, local_rank, px, py, arithmetic, morpho=False, per_rank=1024, steps=5):
    """N > 1: `steps` steps of a per_rank^2-cells-per-GPU problem run decomposed over all ranks against the same
    domain on rank 0's device alone; the gathered blocks must equal the single-device state BIT FOR BIT (the min
    reduction is exact and order-free, everything else is local: SURVEY 8e).  Returns (ok, detail) on rank 0."""
    from kestrel_b200 import capi
    from kestrel_b200.host.settings import Cube
    from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state, rank_block
    T = per_rank // 128

    def runset():
        rs = dambreak_runset(T, 128, morpho=morpho)
        rs.nXtiles, rs.nYtiles = px * T, py * T
        rs.Ytilesize = None
        rs.finalize()
        Lx, Ly = rs.xSize, rs.ySize
        conc = 0.1 if morpho else 0.0
        rs.cubes = [Cube(x=0.0, y=0.0, length=Lx, width=Ly, height=1.0, psi=conc, shape="level"),
                    Cube(x=-0.25 * Lx, y=0.0, length=0.5 * Lx, width=Ly, height=1.0, psi=conc, shape="flat")]
        rs.device = local_rank
        rs.arithmetic = arithmetic
        return rs

    rs = runset()
    rs.comm_rank, rs.comm_size, rs.comm_px, rs.comm_py = rank, world, px, py
    blk = rank_block(rs, rank, px, py)
    q4, b0v = dambreak_state(rs, blk)
    p, keep = rs.to_c()
    st = capi.Stepper(lib, p, keep)
    n = lib.comm_id_bytes()
    idt = torch.zeros(n, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.tensor(list(_new_id(lib)), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    rc = lib.comm_attach(st.h, (capi.C.c_ubyte * n)(*idt.cpu().tolist()))
    assert rc == 0, lib.last_error(st.h)
    st.upload_domain(q4, b0v)
    info = st.integrate_to(1e30, steps)
    mine = torch.from_numpy(st.download_domain()).cuda()
    st.close()
    parts = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, parts, dst=0)
    if rank != 0:
        return None
    rs1 = runset()
    q1, b1 = dambreak_state(rs1)
    p1, keep1 = rs1.to_c()
    s1 = capi.Stepper(lib, p1, keep1)
    s1.upload_domain(q1, b1)
    i1 = s1.integrate_to(1e30, steps)
    ref = s1.download_domain()
    s1.close()
    ok = (info.t, info.nsteps, info.nrefines) == (i1.t, i1.nsteps, i1.nrefines)
    worst = 0.0
    for r in range(world):
        tx0, ty0, ntx, nty = rank_block(rs1, r, px, py)
        sub = ref[:, ty0 * 128:(ty0 + nty) * 128, tx0 * 128:(tx0 + ntx) * 128]
        got = parts[r].cpu().numpy()
        if not np.array_equal(sub, got):
            ok = False
            worst = max(worst, float(np.max(np.abs(sub - got))))
    return ok, {"cells_per_gpu": per_rank * per_rank, "steps": steps, "t": info.t, "max_abs_diff": worst,
                "arithmetic": "contracted" if arithmetic == 1 else "faithful", "workload": "morpho" if morpho else "hydro"}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    cb = cpu_baseline(steps, warm)
    serial = cpu_baseline(steps=2, warmup=1, threads=1)   # Kestrel itself is serial
    cb["serial_value"] = serial["value"]
    line = {"impl": "reference", "metric": "cell-updates/s (fp64)", "value": cb["value"], "unit": "cell-updates/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic dam-break, periodic, xySinSlope(0.2), Chezy 0.04, erosion off (BASELINE.json configs[4])",
                       "cells": CPU_SAMPLE_SIZE ** 2, "note": "bounded sample of the 16384^2 workload; the reference's memory model "
                                                               "(7 KB/cell) cannot hold the full size"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "serial_value")},
            "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--size", type=int, default=0, help="cells per side per GPU (multiple of 128); default 16384")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--arithmetic", type=int, default=1,
                    help="1 (default) contracted fp64, held to the north-star 1e-10 against the oracle; 0 faithful (bit-identical to the oracle)")
    ap.add_argument("--workload", default="hydro", choices=["hydro", "morpho"],
                    help="hydro: the headline dam-break (Chezy, erosion off); morpho: SURVEY 8(d)'s second run (Variable drag, "
                         "Mixed erosion, psi = 0.1), one step = one Strang step H(dt) M(2dt) H(dt)")
    ap.add_argument("--output-intervals", type=int, default=0,
                    help="extra leg: N output intervals of --steps steps each, blocking kgpu_download_domain against the "
                         "asynchronous kgpu_output_begin / kgpu_output_wait (SURVEY 8f rank 2)")
    ap.add_argument("--no-bitwise", action="store_true", help="N > 1: skip the decomposed-vs-single-device bitwise check")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --size cells per side per GPU; strong: --size is the side of the GLOBAL domain, split over the ranks")
    ap.add_argument("--no-faithful", action="store_true", help="skip the side measurement of the faithful-arithmetic variant")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    from kestrel_b200 import capi
    from kestrel_b200.host.synthetic import dambreak_runset, dambreak_state, decomposition, rank_block

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.load_gpu()

    size = args.size or 16384
    free_b, total_b = torch.cuda.mem_get_info()
    morpho = args.workload == "morpho"
    planes = 64 if morpho else 44   # device planes of one handle + staging
    need = planes * (size + 64) ** 2 * 8
    while need > 0.9 * free_b and size > 1024:
        size //= 2
        need = planes * (size + 64) ** 2 * 8
    # host side: every rank pins q4 + out (2 x 32 B/cell) and builds the state with numpy temporaries
    # (~110 B/cell in all); all ranks share one box's RAM, so shrink rather than swap or get killed
    try:
        import psutil
        avail = psutil.virtual_memory().available
        while world * 110 * size * size > 0.7 * avail and size > 1024:
            size //= 2
    except ImportError:
        pass
    if world > 1:  # every rank must agree on the block size
        tsz = torch.tensor([size], device="cuda", dtype=torch.int64)
        dist.all_reduce(tsz, op=dist.ReduceOp.MIN)
        size = int(tsz.item())
    # weak scaling: every GPU owns a size x size block of a (px*size) x (py*size) periodic domain;
    # strong scaling: the size x size domain is split into px x py blocks
    px, py = decomposition(world)
    strong = args.scaling == "strong" and world > 1
    Tx, Ty = (size // 128 // px, size // 128 // py) if strong else (size // 128, size // 128)
    rs = dambreak_runset(Tx, 128, morpho=morpho)
    if world > 1:
        rs.nXtiles, rs.nYtiles = px * Tx, py * Ty
        rs.Ytilesize = None
        rs.finalize()
        from kestrel_b200.host.settings import Cube
        Lx, Ly = rs.xSize, rs.ySize
        conc = 0.1 if morpho else 0.0
        rs.cubes = [Cube(x=0.0, y=0.0, length=Lx, width=Ly, height=1.0, psi=conc, shape="level"),
                    Cube(x=-0.25 * Lx, y=0.0, length=0.5 * Lx, width=Ly, height=1.0, psi=conc, shape="flat")]
        rs.comm_rank, rs.comm_size, rs.comm_px, rs.comm_py = rank, world, px, py
    rs.device = local_rank
    rs.arithmetic = args.arithmetic
    cells = Tx * 128 * Ty * 128
    q4_np, b0v_np = dambreak_state(rs, rank_block(rs, rank, px, py) if world > 1 else None)
    # pinned host buffers (the Fortran host's arrays stand-in) for the e2e leg
    q4 = torch.from_numpy(q4_np).pin_memory()
    b0v = torch.from_numpy(b0v_np).pin_memory()
    del q4_np, b0v_np
    out = torch.empty_like(q4).pin_memory()

    p, keep = rs.to_c()
    st = capi.Stepper(lib, p, keep)
    if world > 1:  # ncclUniqueId from rank 0, broadcast by the host's own means (torch.distributed)
        n = lib.comm_id_bytes()
        idt = torch.zeros(n, dtype=torch.uint8, device="cuda")
        if rank == 0:
            hb = (capi.C.c_ubyte * n)()
            assert lib.comm_create_id(hb) == 0
            idt.copy_(torch.tensor(list(hb), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        idb = (capi.C.c_ubyte * n)(*idt.cpu().tolist())
        rc = lib.comm_attach(st.h, idb)
        assert rc == 0, lib.last_error(st.h)
    st.upload_domain(q4.numpy(), b0v.numpy())
    stream = torch.cuda.ExternalStream(lib.stream(st.h))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    st.integrate_to(1e30, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = lib.launch_count(st.h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    info = st.integrate_to(1e30, args.steps)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = int(lib.launch_count(st.h) - l0)
    nref = int(info.nrefines)
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    value = cells * world * args.steps / (ms * 1e-3)

    # ---- dominant kernel: mean launch duration with CUDA events on the launching stream
    lib.rhs_timing(st.h, None, None, 1)
    st.integrate_to(1e30, 5)
    rms = capi.C.c_double()
    rl = capi.C.c_int64()
    lib.rhs_timing(st.h, capi.C.byref(rms), capi.C.byref(rl), 2)
    k_ms = rms.value / max(rl.value, 1)
    peak, peak_src = load_peaks()
    achieved = cells * ALG_BYTES_PER_CELL_STAGE / (k_ms * 1e-3) / 1e9
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full capture, scaled by cells
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            per_cell = json.load(fh)["dram_bytes_per_cell_per_launch"]["contracted" if args.arithmetic == 1 else "faithful"]
            traffic = per_cell * cells
    except Exception:
        pass
    # second roofline: the fp64 pipe issues one warp instruction per 2 cycles per SM sub-partition;
    # fp64 warp-instructions per cell come from the committed ncu capture, the launch time is live
    fp64 = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            per = json.load(fh)["fp64_pipe_warp_instructions_per_cell_per_launch"]["contracted" if args.arithmetic == 1 else "faithful"]
        props = torch.cuda.get_device_properties(local_rank)
        mhz = clocks.get("sm_mhz") or 1965.0
        pk = props.multi_processor_count * 4 * 0.5 * mhz * 1e6
        ach = per * cells / (k_ms * 1e-3)
        fp64 = {"achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": pk, "frac": ach / pk, "warp_inst_per_cell": per,
                "note": "second roofline (DESIGN.md section 5); peak = SMs x 4 x 0.5 warp-instructions per cycle at the sampled SM clock -- "
                        "measured on this pool's B200: 63.2 DFMA / clk / SM (tools/microbench/fp64_peak.cu, profiles/r02_fp64_peak.json)"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "fp64_pipe": fp64,
                "kernel": "hydro_stage_kernel", "kernel_ms": k_ms, "kernel_launches_timed": int(rl.value),
                "kernel_share_of_step": (8 if morpho else 4) * k_ms / (ms / args.steps), "peak_source": peak_src,
                "step_frac_of_hbm_roofline": cells * (1120 if morpho else ALG_BYTES_PER_CELL_UPDATE) / (ms / args.steps * 1e-3) / 1e9 / peak}

    # ---- end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        k_e2e = args.steps
        barrier()
        t0 = time.perf_counter()
        st.upload_domain(q4.numpy(), b0v.numpy())
        st.integrate_to(1e30, k_e2e)
        st.lib.download_domain(st.h, capi._ptr(out.numpy()), None)
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([t1], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t1 = float(tt.item())
        h2d = (q4.numel() + b0v.numel()) * 8
        d2h = out.numel() * 8
        e2e = {"value": cells * world * k_e2e / t1, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d / k_e2e,
               "d2h_bytes_per_step": d2h / k_e2e, "interval_steps": k_e2e, "interval_s": t1,
               "note": "one output interval: kgpu_upload_domain + kgpu_integrate_to(K steps) + kgpu_download_domain, pinned host buffers"}
    # ---- output intervals: blocking download against the asynchronous gather (host buffers pinned in both)
    pipelined = None
    if args.output_intervals > 0 and world == 1:
        n_int, k_int = args.output_intervals, args.steps
        outs = [out, torch.empty_like(q4).pin_memory()]
        res = {}
        for mode in ("blocking", "async"):
            barrier()
            t0 = time.perf_counter()
            for it in range(n_int):
                st.integrate_to(1e30, k_int)
                if mode == "blocking":
                    st.lib.download_domain(st.h, capi._ptr(outs[it % 2].numpy()), None)
                else:
                    st.output_begin(outs[it % 2].numpy())   # waits for the previous output, snapshots, returns
            if mode == "async":
                st.output_wait()
            torch.cuda.synchronize()
            res[mode] = cells * n_int * k_int / (time.perf_counter() - t0)
        pipelined = {"intervals": n_int, "steps_per_interval": k_int, "unit": "cell-updates/s", "blocking": res["blocking"],
                     "async": res["async"], "d2h_bytes_per_interval": out.numel() * 8,
                     "note": "N x (K steps + whole-state output to pinned host memory); async = kgpu_output_begin/kgpu_output_wait"}
    st.close()

    # ---- the other arithmetic variant beside the headline (device-resident, fewer steps)
    other = None
    if not args.no_faithful:
        rs.arithmetic = 1 - args.arithmetic
        p2, keep2 = rs.to_c()
        st2 = capi.Stepper(lib, p2, keep2)
        if world > 1:
            if rank == 0:
                idt.copy_(torch.tensor(list(_new_id(lib)), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            rc = lib.comm_attach(st2.h, (capi.C.c_ubyte * n)(*idt.cpu().tolist()))
            assert rc == 0, lib.last_error(st2.h)
        st2.upload_domain(q4.numpy(), b0v.numpy())
        stream2 = torch.cuda.ExternalStream(lib.stream(st2.h))
        k2 = max(5, min(args.steps, 20))
        st2.integrate_to(1e30, 3)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream2)
        st2.integrate_to(1e30, k2)
        f1.record(stream2)
        barrier()
        ms2 = f0.elapsed_time(f1)
        if world > 1:
            tms = torch.tensor([ms2], device="cuda", dtype=torch.float64)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms2 = float(tms.item())
        st2.close()
        other = {"arithmetic": "faithful" if args.arithmetic == 1 else "contracted", "value": cells * world * k2 / (ms2 * 1e-3),
                 "unit": "cell-updates/s", "ms_per_step": ms2 / k2, "steps": k2, "warmup": 3}

    # ---- N > 1: the decomposed run equals the single-device run bit for bit (both arithmetic variants)
    decomposed = None
    if world > 1 and not args.no_bitwise:
        checks = [decomposed_bitwise_check(lib, dist, torch, rank, world, local_rank, px, py, a, morpho=morpho) for a in (args.arithmetic, 1 - args.arithmetic)]
        if rank == 0:
            decomposed = {"ok": all(c[0] for c in checks), "runs": [c[1] for c in checks]}

    cb = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cb = cpu_baseline(steps=6, warmup=1, morpho=morpho)
        serial = cpu_baseline(steps=2, warmup=1, threads=1, morpho=morpho)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cb["serial_value"] = serial["value"]  # the Fortran reference is single-threaded

    if rank == 0:
        line = {"metric": "cell-updates/s (fp64)", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": ("synthetic dam-break, periodic, xySinSlope(0.2), Chezy 0.04, erosion off, all tiles active "
                                        "(BASELINE.json configs[4])") if not morpho else
                                       ("synthetic dam-break with morphodynamics (SURVEY 8d second run): Variable drag, Mixed erosion, "
                                        "Spearman-Manning deposition, psi0 = 0.1; one step = one Strang step H(dt) M(2dt) H(dt) = 8 stage "
                                        "launches + 3 morphodynamic stages; 1120 algorithmic B per cell-update"),
                           "cells_per_gpu": cells, "grid": f"{rs.NX}x{rs.NY}", "tiles": f"{rs.nXtiles}x{rs.nYtiles} of 128x128",
                           "decomposition": f"{px}x{py} blocks, 2-cell halos by ncclSend/ncclRecv overlapped with the interior, "
                                            "one ncclAllReduce(min) per dt decision" if world > 1 else "single device",
                           "arithmetic": "faithful fp64 (no FMA contraction, reference operation order; bit-identical to the oracle)" if args.arithmetic == 0
                           else "contracted fp64 (FMA, shared reciprocals; rel-Linf <= 1e-10 per field against the oracle, the north-star bar)",
                           "l2": "fields are 2.1 GB each >> 126 MB L2; no flush needed",
                           "rolled_back_attempts": nref},
                "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
                "other_arithmetic": other}
        if pipelined:
            line["output_intervals"] = pipelined
        if decomposed is not None:
            line["decomposed_bitwise"] = decomposed["ok"]
            line["decomposed_bitwise_runs"] = decomposed["runs"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
