"""Stress of one large registration: repeated reset / buildRBC / run on the same engine, results compared between repetitions."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
m, nr, W, H, reps = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (307200, 512, 640, 480, 30)))
ctx = capi.Context(0)
F = synth.grid_cloud(W, H)
F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
first = None
for outer in range(3):
    s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
    s.init(m, nr, 2e2, 1e-6, W, H)
    s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
    for rep in range(reps):
        s.reset(); s.buildRBC(); s.run(40); ctx.sync()
        T = s.debug("T", np.float32, 8).view(np.uint32)
        if first is None:
            first = T.copy()
        elif not np.array_equal(first, T):
            print("MISMATCH at", outer, rep); sys.exit(1)
    s.close()
print("ok", m, nr, reps * 3, "registrations identical")
