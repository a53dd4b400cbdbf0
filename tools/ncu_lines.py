#!/usr/bin/env python
"""Instruction share by CUDA source line of the first kernel in an ncu report.  Usage: tools/ncu_lines.py rep [N]"""
import csv
import io
import os
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, col, first = "?", None, None
lines = defaultdict(lambda: [0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r[0] == "Function Name":
        if first is None:
            first = r[1]
        elif r[1] != first:
            break
        continue
    if r[0] == "Line No":
        col = {}
        for i, n in enumerate(r):
            col.setdefault(n, i)
        continue
    if col is None or not r[0].isdigit():
        continue
    try:
        s = int(r[col["# Samples"]] or 0)
        n = int(r[col["Instructions Executed"]] or 0)
    except (ValueError, IndexError, KeyError):
        continue
    key = f"{cur_file}:{r[0]} {r[1].strip()[:90]}"
    lines[key][0] += s
    lines[key][1] += n
tot = sum(v[1] for v in lines.values())
print(f"-- instruction share by line (total warp instructions {tot})")
for k, (s, n) in sorted(lines.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"  {100.0 * n / max(1, tot):5.1f}% {n:10d} | {k}")
