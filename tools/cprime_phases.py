"""Phase clocks of kernel C' (k_search_sorted) in the batch engine: set-up + pass 1, item build, item loop, summed over a
40-iteration registration of 64 pairs (ICP_B200_BATCH_PROF attaches the counters)."""
import os, sys
os.environ["ICP_B200_BATCH_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
n, iters = 256, 40
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
b.register(iters); ctx.sync()
tot = np.zeros(8, np.float64)
for p in range(n):
    pr = b.debug("prof", np.uint64, 64, pair=p)
    tot += pr[32:40].astype(np.float64)
ctas = tot[4]
print(f"CTA launches {ctas:.0f}; cycles per CTA: set-up + pass 1 {tot[0]/ctas:.0f}, item build {tot[1]/ctas:.0f}, "
      f"item loop (mean over warps) {tot[2]/ctas/16:.0f}, item loop until the last warp is done {tot[3]/ctas:.0f}; longest CTA of the last pair {tot[5]/n:.0f}; inside pass 1: prologue + prefetch issue {tot[6]/ctas:.0f}, barrier {tot[7]/ctas:.0f}")
