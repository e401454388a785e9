#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and mean time, share.
usage: python tools/ncu_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else ""
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    rows.append((name, r["Grid Size"], r["Block Size"], ns))
agg = OrderedDict()
for name, grid, block, ns in rows:
    k = (name, grid, block)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
total = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary ({len(rows)} launches, {total / 1e6:.3f} ms of kernel time)\n")
if cmd:
    print(f"command: `{cmd}`\n")
print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
print("| kernel | grid | block | launches | total us | mean us | share |")
print("|---|---|---|---:|---:|---:|---:|")
for (name, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {grid} | {block} | {n} | {ns / 1e3:.1f} | {ns / 1e3 / n:.2f} | {100 * ns / total:.1f}% |")
