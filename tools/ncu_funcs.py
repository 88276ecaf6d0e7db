#!/usr/bin/env python
"""Aggregate the CUDA-source view of an `ncu --page source --csv --print-source sass,cuda` dump by the
device FUNCTION each source line belongs to (function spans are read from the current source files).
usage: tools/ncu_funcs.py src.csv [kernel-substring] [column]"""
import csv, re, sys, collections, os
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""
colname = sys.argv[3] if len(sys.argv) > 3 else "Instructions Executed"  # e.g. "# Samples", "stall_long_sb"
rows = list(csv.reader(open(path)))
spans = {}
def func_spans(fpath):
    if fpath in spans: return spans[fpath]
    out = []
    try: lines = open(fpath).read().split("\n")
    except Exception: spans[fpath] = out; return out
    pat = re.compile(r"^\s*(?:template\s*<[^>]*>\s*)?(?:static\s+|__host__\s+|__device__\s+|__global__\s+|__forceinline__\s+|inline\s+|void\s+|const\s+)*[\w:<>,\s\*&]*?\b(\w+)\s*\([^;]*$")
    cur = None
    for i, l in enumerate(lines, 1):
        if ("__device__" in l or "__global__" in l) and "(" in l and not l.strip().startswith("//"):
            m = re.search(r"(\w+)\s*\(", l.split("__device__")[-1] if "__device__" in l else l)
            if m: cur = (m.group(1), i); out.append([m.group(1), i, 10**9])
            if len(out) > 1: out[-2][2] = i - 1
        elif "__global__" in lines[i-2] if i >= 2 else False:
            pass
    spans[fpath] = out; return out
i = 0; fpath = None
tot = collections.Counter(); kern_tot = 0
while i < len(rows):
    r = rows[i]
    if r and r[0] == "File Path": fpath = r[1]
    if r and r[0] == "Function Name":
        fn = r[1]; hdr = rows[i+1]; j = i + 2; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] in ("File Path", "Function Name")):
            body.append(rows[j]); j += 1
        cols = {n: k for k, n in enumerate(hdr)}
        if want in fn and "Line No" in cols and fpath and fpath.endswith((".cu", ".cuh")):
            ie = cols[colname]; ln = cols["Line No"]
            sp = func_spans(fpath)
            for b in body:
                try: n = float(b[ie]); l = int(b[ln])
                except Exception: continue
                name = "?"
                for f, a, z in sp:
                    if a <= l <= z: name = f
                tot[(os.path.basename(fpath), name)] += n
        i = j
    else: i += 1
s = sum(tot.values())
for (f, n), v in tot.most_common(30):
    print(f"{100*v/s:5.1f}%  {f}:{n}")
