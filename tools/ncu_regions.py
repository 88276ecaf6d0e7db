#!/usr/bin/env python
"""Aggregate the CUDA-source view of an `ncu --page source --csv --print-source sass,cuda` dump over line ranges of
one source file.  usage: tools/ncu_regions.py src.csv file-substring  a-b[:label] ...   [--col "Instructions Executed"]"""
import csv, sys
path, fsub = sys.argv[1], sys.argv[2]
col = "Instructions Executed"
regs = []
for a in sys.argv[3:]:
    if a.startswith("--col="): col = a[6:]; continue
    rng, _, lab = a.partition(":")
    lo, hi = rng.split("-")
    regs.append((int(lo), int(hi), lab or rng))
rows = list(csv.reader(open(path)))
i = 0; fpath = None; per_line = {}
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": fpath = r[1]
    if r and r[0] == "Function Name":
        hdr = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
            body.append(rows[j]); j += 1
        cols = {n: k for k, n in enumerate(hdr)}
        if "Line No" in cols and col in cols and fpath and fsub in fpath:
            for b in body:
                try: per_line[int(b[cols["Line No"]])] = per_line.get(int(b[cols["Line No"]]), 0) + float(b[cols[col]])
                except Exception: pass
        i = j
    else: i += 1
tot = sum(per_line.values())
print(f"file *{fsub}*: {col} total {tot:.4g}")
for lo, hi, lab in regs:
    v = sum(n for l, n in per_line.items() if lo <= l <= hi)
    print(f"  {lab:28s} L{lo}-{hi}: {v:.4g}  ({100*v/max(tot,1):.1f}% of file)")
