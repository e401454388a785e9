#!/usr/bin/env python
"""Turn an `ncu --set full` report into a short markdown table + JSON (committed under profiles/).
usage: python tools/ncu_report.py <rep.ncu-rep> <label> [--json profiles/ncu_summary.json --key-suffix _batch --pairs 64]"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__cycles_elapsed.max", "cycles"),
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("label")
ap.add_argument("--json"); ap.add_argument("--key-suffix", default=""); ap.add_argument("--pairs", type=int, default=0)
a = ap.parse_args()
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
agg = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void ", "")
    d = agg.setdefault(name, {"n": 0, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    d["n"] += 1
    for m, k in METRICS:
        if m not in col:
            continue
        try:
            v = float(r[col[m]].replace(",", "")) * SCALE.get(units[col[m]], 1.0)
        except ValueError:
            continue
        d[k] = d.get(k, 0.0) + v
print(f"# ncu --set full summary: {a.label}\n")
print(f"report: `{os.path.basename(a.rep)}` (kept in gpurun_out/, not committed); values are means over the captured launches"
      + (f"; batch of {a.pairs} pairs per launch" if a.pairs else "") + "\n")
print("| kernel | grid | block | n | time us | DRAM rd MB | DRAM wr MB | DRAM % | warp inst | issue % | FMA pipe % | occ % | regs | L1 hit % | L2 hit % | smem conflicts |")
print("|---|---|---|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|--:|")
out = {}
for name, d in agg.items():
    n = d["n"]
    g = lambda k: d.get(k, float("nan")) / n
    print(f"| `{name}` | {d['grid']} | {d['block']} | {n} | {g('time'):.1f} | {g('dram_rd') / 1e6:.2f} | {g('dram_wr') / 1e6:.2f} | {g('dram_pct'):.1f} | "
          f"{g('warp_inst'):.3g} | {g('issue_pct'):.1f} | {g('fma_pipe_pct'):.1f} | {g('occupancy_pct'):.1f} | {g('regs'):.0f} | {g('l1_hit_pct'):.1f} | "
          f"{g('l2_hit_pct'):.1f} | {g('smem_conflicts'):.0f} |")
    key = re.sub(r"<.*", "", name) + a.key_suffix
    out[key] = {"kernel": name, "grid": d["grid"], "block": d["block"], "pairs_per_launch": a.pairs or None,
                "time_us_under_ncu": g("time"), "dram_bytes_per_launch": g("dram_rd") + g("dram_wr"),
                "warp_inst_per_launch": g("warp_inst"), "issue_active_pct": g("issue_pct"), "fma_pipe_pct": g("fma_pipe_pct"),
                "occupancy_pct": g("occupancy_pct"), "registers": g("regs")}
if a.json:
    cur = {}
    if os.path.exists(a.json):
        cur = json.load(open(a.json))
    cur.update(out)
    json.dump(cur, open(a.json, "w"), indent=1)
