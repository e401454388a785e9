"""Small batch-mode + latency-mode run for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
ctx = capi.Context(0)
base = ctx.upload(synth.base_landmarks())
b = alg.ICPBatch(ctx, 10, 16384, 256)
b.synthesize(base, 5000)
b.set_slices(2)
b.register(4)
print("batch poses finite:", np.isfinite(b.read_poses()).all(), "cmode", b.cmode())
b.close()
F, M, _, _ = synth.known_transform_pair(seed=42)
for rot in (1, 0):
    s = alg.ICPStep(ctx, rot, 1); s.init(16384, 256, 2e2, 1e-6); s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
    s.buildRBC(); s.run(3); ctx.sync()
    print("single pose", s.debug("T", np.float32, 8))
    s.close()
# one larger image-ordered registration: the wide kernel D (whole-GPU launches per pass) and kernel C over 32 x 16 patches
os.environ["ICP_B200_WIDED"] = "1"; os.environ["ICP_B200_CTILE_WH"] = "32x16"
Fg = synth.grid_cloud(256, 256)
F2, M2, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=Fg)
s = alg.ICPStep(ctx, 1, 1); s.init(65536, 512, 2e2, 1e-6, 256, 256); s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M2)
s.buildRBC(); s.run(2); ctx.sync()
print("large single pose", s.debug("T", np.float32, 8))
s.close()
