"""Small profiling target for ncu: a few fused iterations of one pair (latency mode) and of a batch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth

what = sys.argv[1] if len(sys.argv) > 1 else "single"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rot = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = capi.Context(0)
F, M, _, _ = synth.known_transform_pair(seed=42)
if what in ("single", "staged"):
    s = alg.ICPStep(ctx, rot, 1); s.init(16384, 256, 2e2, 1e-6)
    s.set_mode(0 if what == "staged" else 1)
    s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
    s.buildRBC(); s.run(iters, variant=0); ctx.sync()
    print("T", s.debug("T", np.float32, 8))
else:
    n_pairs = int(what)
    b = alg.ICPBatch(ctx, n_pairs, 16384, 256, rot=rot)
    base = ctx.upload(synth.base_landmarks())
    b.synthesize(base, 5000)
    b.register(iters); ctx.sync()
    print("T0", b.read_poses()[0])
