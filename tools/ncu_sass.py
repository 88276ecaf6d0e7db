#!/usr/bin/env python
"""SASS rows (with their CUDA line and per-datapoint execution counts) of one source-line range, from an
`ncu --page source --csv --print-source sass,cuda` dump.
usage: tools/ncu_sass.py src.csv file-substring lo hi datapoints_per_launch"""
import csv, sys
path, fsub, lo, hi, N = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
rows = list(csv.reader(open(path)))
i = 0; fpath = None; done = False
while i < len(rows) and not done:
    r = rows[i]
    if r and r[0] == "File Path": fpath = r[1]
    if r and r[0] == "Function Name":
        hdr = rows[i + 1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
            body.append(rows[j]); j += 1
        if fpath and fsub in fpath and "Address" in hdr:
            ln, ad = hdr.index("Line No"), hdr.index("Address")
            src = [k for k, n in enumerate(hdr) if n == "Source"][-1]
            ie = hdr.index("Instructions Executed")
            tot = 0.0
            for b in body:
                if len(b) > ie and b[ad] and b[ln].isdigit() and lo <= int(b[ln]) <= hi:
                    n = float(b[ie] or 0) / N
                    tot += n
                    print(f"{n:7.2f} L{b[ln]:>4} {b[src][:100]}")
            print(f"total {tot:.1f} warp instructions per datapoint in L{lo}-{hi}")
            done = True
        i = j
    else: i += 1
