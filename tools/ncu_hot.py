#!/usr/bin/env python
"""Stall-reason and opcode summary + hottest SASS lines of one kernel from an ncu report (source page).
usage: python tools/ncu_hot.py <rep> <kernel regex> [top N]"""
import csv, io, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1])
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ("Kernel Name", "Address"): break
    data.append(r)
S = lambda r, k: int(r[col[k]] or 0)
tot_s = sum(S(r, "# Samples") for r in data); tot_i = sum(S(r, "Instructions Executed") for r in data)
print(f"{len(data)} SASS lines, {tot_s} samples, {tot_i} warp instructions")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(S(r, s) for r in data) for s in stalls}
print("stalls:", ", ".join(f"{s[6:]} {100 * v / max(tot_s, 1):.1f}%" for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
mix = Counter()
for r in data:
    t = r[col["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    mix[op.split(".")[0]] += S(r, "Instructions Executed")
print("opcodes:", ", ".join(f"{o} {100 * v / max(tot_i, 1):.1f}%" for o, v in mix.most_common(14)))
print("hottest lines (samples, executed, top stall, SASS):")
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:top]:
    st = max(stalls, key=lambda s: S(r, s))
    print(f"  {S(r, '# Samples'):6d} {S(r, 'Instructions Executed'):10d} {st[6:]:14s} {r[col['Source']].strip()[:90]}")
