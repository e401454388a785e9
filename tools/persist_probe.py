import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
ctx = capi.Context(0)
F, M, _, _ = synth.batch_pair(9000)
s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
s.init(16384, 256, 2e2, 1e-6)
s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
s.buildRBC(); ctx.sync()
s.run(int(sys.argv[1]) if len(sys.argv) > 1 else 1, variant=3); ctx.sync()
print("T", s.debug("T", np.float32, 8))
