#!/usr/bin/env python
"""Turn one tools/profile_cmd.sh run (gpurun_out/{launches,prof,plain}_<tag>.*) into the tracked
profiles/<name>_launches.csv + profiles/<name>_summary.md.

usage: tools/profile_summary.py <tag> <name> ["title"]      e.g.  r1d r01d "Gram-form E-step"
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, name = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else ""
out = os.path.join(ROOT, "gpurun_out")
prof = os.path.join(ROOT, "profiles")

# ---- launch list ---------------------------------------------------------------------------------
src = os.path.join(out, f"launches_{tag}.csv")
shutil.copy(src, os.path.join(prof, f"{name}_launches.csv"))
rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t, n = collections.Counter(), collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
for r in rows[1:]:
    if len(r) <= vi:
        continue
    k = r[ki][:70]
    t[k] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1e-6)
    n[k] += 1
tot = sum(t.values())

# ---- full-section capture ------------------------------------------------------------------------
rep = os.path.join(out, f"prof_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
want = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"),
    ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
]
stalls = [(i, c.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for i, c in enumerate(h) if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]

lines = [f"# {name}: {title}", "",
         "Command (under gpurun, one B200): `python bench.py --n 131072 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other`",
         "(`tools/profile_cmd.sh`): plain run first, then `ncu --metrics gpu__time_duration.sum --clock-control none -c 400`",
         f"(launch list -> `profiles/{name}_launches.csv`), then `ncu --set full --clock-control none --import-source on`",
         "on the E-step / M-step kernels.  Times under ncu are serialised and cold-cache; the bench numbers come from the plain run.",
         ""]
try:
    plain = [l for l in open(os.path.join(out, f"plain_{tag}.log")) if l.startswith("{")]
    j = json.loads(plain[-1])
    lines += [f"Plain run: E-step {j['estep_ms']:.3f} ms for {j['config']['N_per_gpu']} datapoints = "
              f"{j['value'] / 1e6:.2f} M datapoints/s; EM iteration {j['em_iter_ms']:.3f} ms.", ""]
except Exception:
    pass
lines += ["## Launch list (share of GPU time, warm-up + timed steps)", "", "| kernel | launches | ms | share |", "|---|---|---|---|"]
for k, v in t.most_common(10):
    lines.append(f"| `{k}` | {n[k]} | {v:.3f} | {100 * v / tot:.1f}% |")
lines += ["", "## `ncu --set full`, per captured launch", ""]
caps = rr[2:]
kn = h.index("Kernel Name")
lines.append("| metric | " + " | ".join(f"`{c[kn].split('(')[0][-40:]}`" for c in caps) + " |")
lines.append("|---|" + "---|" * len(caps))
for key, label in want:
    if key in h:
        i = h.index(key)
        lines.append(f"| {label} ({rr[1][i]}) | " + " | ".join(c[i] for c in caps) + " |")
lines += ["", "Warp stall reasons (warps stalled per issue-active cycle; the largest are what the kernel waits on):", ""]
lines.append("| stall | " + " | ".join(str(k) for k in range(len(caps))) + " |")
lines.append("|---|" + "---|" * len(caps))
st = sorted(stalls, key=lambda s: -float(caps[0][s[0]] or 0))[:8]
for i, label in st:
    lines.append(f"| {label} | " + " | ".join(c[i] for c in caps) + " |")

# ---- per-function instruction share --------------------------------------------------------------
srcp = f"/tmp/{tag}_src.csv"
with open(srcp, "w") as f:
    f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                           capture_output=True, text=True).stdout)
fn = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_funcs.py"), srcp, "estep"], capture_output=True, text=True).stdout
lines += ["", "## E-step kernel: share of executed warp instructions per device function (ncu source page, -lineinfo)", "", "```"]
lines += fn.splitlines()[:16] + ["```", ""]
open(os.path.join(prof, f"{name}_summary.md"), "w").write("\n".join(lines))
print("\n".join(lines))
