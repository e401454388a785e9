#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py <rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern,
                      "--launch-count", "1", "--launch-skip", __import__("os").environ.get("SKIP", "0")], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
tot = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or r[0] in ("Function Name",):
        continue
    if r[0].strip().isdigit():
        try:
            inst = int(r[hdr["Instructions Executed"]]); samp = int(r[hdr["# Samples"]])
        except (ValueError, KeyError):
            continue
        key = (cur_file, int(r[0]), r[1].strip()[:110])
        a = tot.setdefault(key, [0, 0]); a[0] += inst; a[1] += samp
ti = sum(v[0] for v in tot.values()) or 1; ts = sum(v[1] for v in tot.values()) or 1
print(f"total warp instructions {ti}, samples {ts}")
for (f, ln, src), (i, s) in sorted(tot.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("BY_INST","0")) ^ 1])[:top]:
    print(f"{100*s/ts:5.1f}% samp {100*i/ti:5.1f}% inst  {f}:{ln}  {src}")
