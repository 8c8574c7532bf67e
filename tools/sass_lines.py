#!/usr/bin/env python
"""Static SASS histogram per source line for one kernel of a cubin (needs -lineinfo).

usage: sass_lines.py <cubin> <mangled-kernel-substring> [top]
Extract cubins with `cuobjdump -xelf all kestrel_b200/lib/libkestrel_gpu.so`.
"""
import collections
import re
import subprocess
import sys

cubin, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
on = False
cur = None
cnt = collections.Counter()
ops = collections.defaultdict(collections.Counter)
allops = collections.Counter()
for l in txt:
    m = re.match(r"\s+\.global\s+(\S+)", l)
    if m:
        on = pat in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        op = m.group(2).split(".")[0]
        cnt[cur] += 1
        ops[cur][op] += 1
        allops[op] += 1
tot = sum(cnt.values())
print("total instructions", tot)
print(dict(allops.most_common(25)))
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    print(f"{k[0]}:{k[1]:<5d} {v:5d}  {dict(ops[k].most_common(6))}")
