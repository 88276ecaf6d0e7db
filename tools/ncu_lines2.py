#!/usr/bin/env python
"""Hottest source lines (warp-stall samples) of an `ncu --page source --csv --print-source sass,cuda` dump.
usage: tools/ncu_lines2.py src.csv [top]"""
import collections
import csv
import glob
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
i, fpath = 0, None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0.0])
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path":
        fpath = r[1]
    if r and r[0] == "Function Name":
        hdr = rows[i + 1]
        j = i + 2
        if "Line No" in hdr:
            c = {n: k for k, n in enumerate(hdr)}
            while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
                b = rows[j]
                try:
                    a = agg[(os.path.basename(fpath), int(b[c["Line No"]]))]
                    a[0] += float(b[c["# Samples"]])
                    a[1] += float(b[c["stall_long_sb"]])
                    a[2] += float(b[c["Instructions Executed"]])
                    a[3] += float(b[c["stall_short_sb"]])
                    a[4] += float(b[c["stall_wait"]])
                except Exception:
                    pass
                j += 1
        i = j
        continue
    i += 1
tot = sum(a[0] for a in agg.values())
src = {}
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tvo_b200", "csrc")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        p = glob.glob(os.path.join(root, f))
        src[f] = open(p[0]).read().split("\n") if p else []
    line = src[f][l - 1].strip()[:80] if src[f] and l <= len(src[f]) else ""
    print(f"{100 * a[0] / tot:5.1f}%  long_sb {a[1]:7.0f}  short_sb {a[3]:7.0f}  wait {a[4]:6.0f}  inst {a[2] / 1e6:7.1f}M  {f}:{l}  {line}")
