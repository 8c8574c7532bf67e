#!/usr/bin/env python
"""Dynamic instruction mix of one profiled kernel from an ncu report (needs --import-source on).

usage: ncu_opmix.py <report.ncu-rep> [cells]
Prints executed warp-instructions per opcode, and (cuda,sass view) per source line.
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
cells = float(sys.argv[2]) if len(sys.argv) > 2 else None


def page(view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
    lines = out.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"') or l.startswith('"#"') or l.startswith('"Line'))
    return list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))


rows = page("sass")
ops = collections.Counter()
samples = collections.Counter()
tot = 0
for r in rows:
    src = r["Source"].strip()
    toks = src.split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0]
    n = int(r["Instructions Executed"] or 0)
    ops[op] += n
    samples[op] += int(r["# Samples"] or 0)
    tot += n
print("warp instructions executed:", tot, "" if not cells else f"= {tot / cells:.1f} per cell")
ts = sum(samples.values())
for op, n in ops.most_common(30):
    print(f"{op:10s} {n:12d} {100.0 * n / tot:5.1f}%   stall samples {100.0 * samples[op] / max(ts, 1):5.1f}%")
try:
    rows = page("cuda,sass")
except Exception as e:  # pragma: no cover
    print("no cuda view:", e)
    sys.exit(0)
