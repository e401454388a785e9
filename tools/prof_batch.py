"""ncu target: every fused kernel launched in-stream (graphs with conditional nodes are not profilable) on a 64-pair batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n_pairs, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
b.register(3); ctx.sync()
for which in range(4):
    print(which, b.time_kernel(which, 2))
