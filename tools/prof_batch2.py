"""ncu target: a 64-pair batch registered in-stream (ICP_B200_NO_GRAPH=1: every fused kernel is a plain launch), 14
iterations, so that late launches see the temporal pruning at work.  usage: python tools/prof_batch2.py [pairs] [iters]"""
import os
import sys

os.environ["ICP_B200_NO_GRAPH"] = "1"
os.environ.setdefault("ICP_B200_BATCH_SLICES", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 14
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, n_pairs, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
b.register(iters); ctx.sync()
print("cmode", b.cmode(), b.config())
