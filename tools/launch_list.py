"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` log into the per-kernel launch list kept
under profiles/ (mean time per launch and share of a step; every kernel of the step launches once
per step, so share = its mean / the sum of the means).  Usage: launch_list.py <csv>"""
import csv, sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
t = defaultdict(list)
for r in rows:
    if r is hdr or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    t[r[ki]].append(v)
# bank construction kernels are not part of a step
step_kernels = {k: v for k, v in t.items() if "pack_rows" not in k and "sample_patches" not in k and k.startswith(("void hb::", "hb::"))}
step_us = sum(sum(v) / len(v) for v in step_kernels.values())
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    if not k.startswith(("void hb::", "hb::")):
        continue
    share = (sum(v) / len(v)) / step_us if k in step_kernels else float("nan")
    print(f"{k[:96]:96s} launches={len(v):3d} mean_us={sum(v) / len(v):10.1f} share_of_step={share:.4f}")
