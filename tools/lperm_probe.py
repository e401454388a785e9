import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
n, nr, it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
m = 16384
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n, m, nr)
QB = b.config()["QB"]
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
b.register(it); ctx.sync()
bad = 0
for p in range(n):
    nbx = b.debug("nbx", np.uint32, m * 16, pair=p)
    lperm = nbx[2 * m: 2 * m + m // 2].view(np.uint16)
    q_rep = b.debug("q_rep", np.uint32, m, pair=p)
    for c in range(m // QB):
        seg = lperm[c * QB:(c + 1) * QB]
        if not np.array_equal(np.sort(seg), np.arange(QB, dtype=np.uint16)):
            bad += 1
            if bad <= 3:
                u, cnt = np.unique(seg, return_counts=True)
                print("pair", p, "chunk", c, "dups", u[cnt > 1][:8], "max", seg.max(), "missing", np.setdiff1d(np.arange(QB), seg)[:8])
print("QB", QB, "bad chunks", bad, "wconst", b.debug("wconst", np.uint32, 16, pair=0)[[0, 1, 2, 12, 13]])
