"""ncu target: ONE registration at a scaled size (default 307200 landmarks / 1024 representatives), in-stream launches
(variant 0), 12 iterations so that late launches see the temporal pruning.  usage: python tools/prof_scaled.py [m nr W H iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth
m, nr, W, H, iters = (int(x) for x in (sys.argv[1:6] if len(sys.argv) >= 6 else (307200, 1024, 640, 480, 12)))
ctx = capi.Context(0)
F = synth.grid_cloud(W, H)
F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
s.init(m, nr, 2e2, 1e-6, W, H)
s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
s.buildRBC(); ctx.sync()
s.run(iters, variant=0); ctx.sync()
print("done", m, nr)
