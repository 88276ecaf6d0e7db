#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line.
usage: tools/ncu_lines.py src.csv [top]"""
import csv, sys, collections, re
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Function Name":
        fn = rows[i][1][:80]
        hdr = rows[i + 1]
        # sass view first: rows until next "File Path"/"Function Name"
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
            body.append(rows[j]); j += 1
        cols = {n: k for k, n in enumerate(hdr)}
        if "Source" in cols and "Instructions Executed" in cols and "Line No" in cols:
            ie = cols["Instructions Executed"]; ss = cols.get("# Samples")
            tot = sum(float(r[ie] or 0) for r in body if len(r) > ie and r[ie].replace('.','').isdigit())
            tots = sum(float(r[ss] or 0) for r in body if ss is not None and len(r) > ss and r[ss].replace('.','').isdigit())
            print(f"\n== {fn}\n   view with 'Line No': {len(body)} rows, instr {tot:.3e}, samples {tots:.0f}")
            agg = []
            for r in body:
                try: agg.append((float(r[ie] or 0), float(r[ss] or 0) if ss is not None else 0, r[cols["Line No"]], r[1][:110]))
                except Exception: pass
            for n, s, ln, src in sorted(agg, reverse=True)[:top]:
                print(f"   {100*n/max(tot,1):5.1f}% instr {100*s/max(tots,1):5.1f}% smp  L{ln:>4}  {src}")
        i = j
    else:
        i += 1
