"""Phase clocks of kernel A (block 0 of every pair, last iteration of the registration) in the batch engine."""
import os, sys
os.environ["ICP_B200_BATCH_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
n = 256
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
for iters in (2, 10, 40):
    b.register(iters); ctx.sync()
    ph = np.zeros(4); nf = np.zeros(2)
    for p in range(n):
        pr = b.debug("prof", np.uint64, 64, pair=p).astype(np.int64)
        nf += pr[44:46]
        c = pr[16 + 1:16 + 6]
        ph += np.diff(c)[:4]
    ph /= n
    print(f"iteration {iters}: stage reps {ph[0]:.0f}, pruned pass {ph[1]:.0f}, full scan {ph[2]:.0f}, rank + store {ph[3]:.0f} cycles (block 0, mean over pairs); fallback points per chunk, mean over the whole registration {nf[0] / max(nf[1], 1):.1f}")
