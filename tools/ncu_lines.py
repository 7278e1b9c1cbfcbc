#!/usr/bin/env python
"""Attribute executed warp-instructions of one kernel (from an .ncu-rep) to CUDA source lines.
usage: ncu_lines.py rep kernel_regex cubin mangled_substring [launch_index]"""
import csv, re, subprocess, sys, collections
rep, kre, cubin, mangled = sys.argv[1:5]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
tables, cur = [], None
for r in rows:
    if 'Address' in r and 'Source' in r:
        cur = {"hdr": r, "rows": []}; tables.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
t = tables[which]
ie = t["hdr"].index("Instructions Executed")
execs = [float(r[ie] or 0) for r in t["rows"]]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".text.") and mangled in l)
cur_line = None; per = []
for l in lines[start + 1:]:
    if l.strip().startswith(".text.") or l.strip().startswith(".section"):
        if per: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per.append(cur_line)
n = min(len(per), len(execs))
agg = collections.Counter()
for i in range(n):
    agg[per[i]] += execs[i]
tot = sum(execs)
print(f"instructions: sass rows {len(execs)} disasm {len(per)} total executed {tot:.3g}")
for (k, v) in agg.most_common(30):
    print(f"{v / tot * 100:5.1f}%  {k}")
