import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
it = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 1):
    b.register(it); ctx.sync()
print("ok", b.read_poses()[:2])
