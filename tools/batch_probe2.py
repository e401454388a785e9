import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import pyoracle as po
from icp_b200 import algorithms as alg, capi, synth
n, nr, it = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n, 16384, nr)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
b.register(it); ctx.sync()
T8 = b.read_poses()
po.set_threads(16)
bad = 0
for p in (0, n // 2, n - 1):
    F = b.debug("F", np.float32, (16384, 8), pair=p); M_ = b.debug("M", np.float32, (16384, 8), pair=p)
    ref = po.icp_register(F, M_, 128, 128, nr, fixed_iters=it)
    bad += not np.array_equal(ref["T"].view(np.uint32), T8[p].view(np.uint32))
print("pairs", n, "nr", nr, "mismatches", bad, b.config())
