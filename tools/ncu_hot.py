#!/usr/bin/env python
"""Summarise an ncu report's source page: hottest CUDA source lines (by stall samples and by
executed instructions) and the stall-reason mix.  Usage: tools/ncu_hot.py report.ncu-rep [N]"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, col, h = "?", None, None
lines = defaultdict(lambda: [0, 0])
stall = defaultdict(int)
first_fn = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r[0] == "Function Name":
        if first_fn is None:
            first_fn = r[1]
        elif r[1] != first_fn:
            break           # only the first kernel instance
        continue
    if r[0] == "Line No":
        h = r
        col = {}
        for i, n in enumerate(h):
            col.setdefault(n, i)
        continue
    if col is None or not r[0].isdigit():
        continue
    try:
        s = int(r[col["# Samples"]] or 0)
        n = int(r[col["Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    key = f"{cur_file}:{r[0]} {r[1].strip()[:95]}"
    lines[key][0] += s
    lines[key][1] += n
    for name, i in col.items():
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                stall[name] += int(r[i] or 0)
            except (ValueError, IndexError):
                pass
tot_s = sum(v[0] for v in lines.values())
tot_i = sum(v[1] for v in lines.values())
print(f"{first_fn[:90]}\ntotal samples {tot_s}, warp instructions {tot_i}")
print("-- stall mix")
for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {100.0 * v / max(1, sum(stall.values())):5.1f}%")
print("-- top lines by samples (share of samples | share of instructions)")
for k, (s, n) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"  {100.0 * s / max(1, tot_s):5.1f}% {100.0 * n / max(1, tot_i):5.1f}% | {k}")
