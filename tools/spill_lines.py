#!/usr/bin/env python
"""Where does a kernel spill?  STL / LDL instructions of one kernel of an object file, by source line (needs -lineinfo).
usage: python tools/spill_lines.py icp_b200/csrc/icp_fused.o _Z15k_search_sortedILb1"""
import re, subprocess, sys, tempfile, os, glob
from collections import Counter
obj, sym = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = glob.glob(d + "/*.cubin")[0]
sass = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text." + sym))
end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.")), len(sass))
cur = None; stl = Counter(); ldl = Counter(); n = 0
for l in sass[start:end]:
    m = re.search(r'//## File ".*?([\w.]+)", line (\d+)', l)
    if m: cur = (m.group(1), int(m.group(2))); continue
    if re.search(r'^\s+/\*[0-9a-f]+\*/', l):
        n += 1
        if " STL" in l: stl[cur] += 1
        if " LDL" in l: ldl[cur] += 1
print("instructions", n)
print("STL", sorted(stl.items()))
print("LDL", sorted(ldl.items()))
