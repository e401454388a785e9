"""Phase clocks of kernel D's body in the batch engine (the tail of k_search_sorted<true>, one CTA per pair): cycles of the last
iteration of a 40-iteration registration of 256 pairs, mean over the pairs (ICP_B200_BATCH_PROF attaches the counters)."""
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
d = np.zeros(6, np.float64)
for p in range(n):
    pr = b.debug("prof", np.uint64, 64, pair=p).astype(np.int64)
    t = [pr[0], pr[1], pr[2], pr[3], pr[4], pr[6], pr[5]]
    d += np.diff(np.array(t, np.float64))
d /= n
print(f"kernel D tail, cycles per pair (last iteration): sum_w {d[0]:.0f}, means {d[1]:.0f}, S partials {d[2]:.0f}, S finish {d[3]:.0f}, "
      f"power method {d[4]:.0f}, accumulate + loop control {d[5]:.0f}; total {d.sum():.0f}")
