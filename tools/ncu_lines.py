#!/usr/bin/env python
"""Dynamic instruction counts per source line: ncu SASS page zipped with nvdisasm -g line info.

usage: ncu_lines.py <report.ncu-rep> <cubin of the same build> <mangled-kernel-substring> [cells] [top]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, pat = sys.argv[1:4]
cells = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
# a report with several captured launches prints one table per launch: KERNEL_INDEX (env, default 0) picks one
import os
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')]
ki = int(os.environ.get("KERNEL_INDEX", "0"))
start = starts[ki]
stop = len(lines)
for j in range(start + 1, len(lines)):
    if not lines[j].startswith('"0x'):
        stop = j
        break
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:stop]))))

txt = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
on = False
cur = None     # innermost source line of the instruction
outer = None   # line of the kernel body it was inlined into (end of the "inlined at" chain)
chain = False
info = []
for l in txt:
    m = re.match(r"\s+\.global\s+(\S+)", l)
    if m:
        on = pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        here = (m.group(1).split("/")[-1], int(m.group(2)))
        if not chain:
            cur = here
        outer = here
        chain = True
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m:
        info.append((cur, outer, m.group(2).split(".")[0]))
    chain = False
print("ncu rows", len(rows), "nvdisasm instrs", len(info))
n = min(len(rows), len(info))
byline = collections.Counter()
byouter = collections.Counter()
opsouter = collections.defaultdict(collections.Counter)
ops = collections.defaultdict(collections.Counter)
stalls = collections.Counter()
tot = 0
mism = 0
for r, (cur, stack, op) in zip(rows[:n], info[:n]):
    toks = r["Source"].split()
    rop = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    if rop != op:
        mism += 1
    k = int(r["Instructions Executed"] or 0)
    tot += k
    byline[cur] += k
    ops[cur][op] += k
    stalls[cur] += int(r["# Samples"] or 0)
    byouter[stack or cur] += k
    opsouter[stack or cur][op] += k
print("opcode mismatches", mism, " total warp instr", tot, f"= {tot / cells:.2f} per cell")
ts = sum(stalls.values())
print("--- by innermost line")
for k, v in byline.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} {v / cells:7.3f}/cell {100.0 * v / tot:5.1f}%  stall {100.0 * stalls[k] / max(ts, 1):5.1f}%  {dict((o, round(c / cells, 2)) for o, c in ops[k].most_common(5))}")
print("--- by line of the kernel body (inlined callees attributed to their call site)")
for k, v in byouter.most_common(top):
    print(f"{k[0]}:{k[1]:<5d} {v / cells:7.3f}/cell {100.0 * v / tot:5.1f}%  {dict((o, round(c / cells, 2)) for o, c in opsouter[k].most_common(5))}")
import os
rng = os.environ.get("PHASES")  # e.g. "283:stage,312:A,343:C,515:D,652:cfl,678:end"
if rng:
    marks = [(int(a), b) for a, b in (x.split(":") for x in rng.split(","))]
    agg = collections.Counter()
    aggops = collections.defaultdict(collections.Counter)
    for (f, ln), v in byouter.items():
        name = "pre"
        for a, b in marks:
            if ln >= a:
                name = b
        agg[name] += v
        aggops[name].update(opsouter[(f, ln)])
    stallcols = [c for c in rows[0].keys() if c.startswith("stall_") and not c.endswith("(Not Issued)")]
    phase_stall = collections.defaultdict(collections.Counter)
    phase_samples = collections.Counter()
    for r, (cur, stack, op) in zip(rows[:n], info[:n]):
        ln = (stack or cur)[1] if (stack or cur) else 0
        name = "pre"
        for a, b in marks:
            if ln >= a:
                name = b
        phase_samples[name] += int(r["# Samples"] or 0)
        for c in stallcols:
            phase_stall[name][c] += int(r[c] or 0)
    tsamp = sum(phase_samples.values())
    print("--- warp-state samples by phase (share of all samples; top stall reasons within the phase)")
    for k, v in phase_samples.most_common():
        top5 = ", ".join(f"{c[6:]} {100.0 * x / max(v, 1):.0f}%" for c, x in phase_stall[k].most_common(6))
        print(f"{k:8s} {100.0 * v / max(tsamp, 1):5.1f}% of samples: {top5}")
    print("--- by phase")
    for k, v in agg.most_common():
        print(f"{k:8s} {v / cells:7.3f}/cell {100.0 * v / tot:5.1f}%  {dict((o, round(c / cells, 2)) for o, c in aggops[k].most_common(14))}")
