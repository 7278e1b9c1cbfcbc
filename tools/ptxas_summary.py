#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel family with max registers / spills / smem."""
import re, sys, subprocess, collections
txt = sys.stdin.read()
cur = None
rows = collections.defaultdict(lambda: dict(n=0, regs=0, spill=0, smem=0, stack=0))
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = m.group(1)
        d = re.search(r"_ZN4ibvh\d+([a-z_0-9]+?)I", name)
        cur = d.group(1) if d else name[:40]
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
    if m and cur:
        rows[cur]["stack"] = max(rows[cur]["stack"], int(m.group(1))); rows[cur]["spill"] = max(rows[cur]["spill"], int(m.group(2)))
    m = re.search(r"Used (\d+) registers", line)
    if m and cur:
        r = rows[cur]; r["n"] += 1; r["regs"] = max(r["regs"], int(m.group(1)))
        s = re.search(r"(\d+) bytes smem", line)
        if s: r["smem"] = max(r["smem"], int(s.group(1)))
for k, r in sorted(rows.items()):
    print(f"{k:28s} inst={r['n']:3d} max_regs={r['regs']:3d} stack={r['stack']:4d} spill={r['spill']:4d} smem={r['smem']}")
