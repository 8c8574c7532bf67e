#!/usr/bin/env python
"""Condense an ncu --set full report of the stage kernel into a small JSON summary for profiles/.

usage: ncu_summary.py <report.ncu-rep> <cells in the capture> <out.json> [note]
"""
import collections
import csv
import io
import json
import subprocess
import sys

rep, cells, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
note = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# a report with several captured launches has one value row per launch: KERNEL_INDEX (env, default: the last) picks one
import os
_ki = os.environ.get("KERNEL_INDEX")
hdr, units, val = rows[0], rows[1], (rows[2 + int(_ki)] if _ki is not None else rows[-1])
get = lambda k: (val[hdr.index(k)], units[hdr.index(k)]) if k in hdr else None
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
summ = {"kernel": val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", "note": note, "cells": cells, "metrics": {}}
for k in keys:
    g = get(k)
    if g:
        summ["metrics"][k] = {"value": g[0], "unit": g[1]}
stalls = {}
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(val[i])
summ["stall_cycles_per_issued_instruction"] = dict(sorted(stalls.items(), key=lambda x: -x[1])[:10])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(src) if l.startswith('"Address"')]
start = starts[int(_ki)] if _ki is not None else starts[-1]
stop = len(src)
for j in range(start + 1, len(src)):
    if not src[j].startswith('"0x'):
        stop = j
        break
ops = collections.Counter()
for r in csv.DictReader(io.StringIO("\n".join(src[start:stop]))):
    t = r["Source"].split()
    if not t:
        continue
    op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    ops[op] += int(r["Instructions Executed"] or 0)
tot = sum(ops.values())
summ["warp_instructions_per_cell"] = tot / cells
summ["opcode_mix_pct"] = {k: round(100.0 * v / tot, 2) for k, v in ops.most_common(16)}
mt = summ["metrics"]
try:
    rd, wr = float(mt["dram__bytes_read.sum"]["value"]), float(mt["dram__bytes_write.sum"]["value"])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    b = rd * scale[mt["dram__bytes_read.sum"]["unit"]] + wr * scale[mt["dram__bytes_write.sum"]["unit"]]
    summ["dram_bytes_per_cell"] = b / cells
    t_us = float(mt["gpu__time_duration.sum"]["value"]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[mt["gpu__time_duration.sum"]["unit"].replace("second", "s").replace("usecond", "us")]
    summ["dram_GBps_under_profiler"] = b / (t_us * 1e-6) / 1e9
except Exception as e:  # pragma: no cover
    summ["dram_note"] = str(e)
json.dump(summ, open(out, "w"), indent=1)
print(json.dumps(summ, indent=1)[:1500])
