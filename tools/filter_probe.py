"""Candidate filter of the stage-2 list scans: per iteration, how many list points are looked at geometry-only, how many are
fully evaluated (unfiltered tiles + candidates), against the points visited.  24 pairs of the bench workload."""
import os, sys
os.environ["ICP_B200_BATCH_EVALS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
n = 24
ctx = capi.Context(0)
base = ctx.upload(synth.base_landmarks())
prev = np.zeros(6)
rows = []
for K in (1, 2, 3, 5, 8, 12, 16, 20, 25, 30, 35, 40):
    b = alg.ICPBatch(ctx, n, 16384, 256)
    b.synthesize(base, 5000)
    b.register(K); ctx.sync()
    ev = np.stack([b.debug("evals", np.uint64, 6, pair=p) for p in range(n)]).astype(np.float64).sum(0) / n
    rows.append((K, ev.copy()))
    b.close()
pk, pe = 0, np.zeros(6)
for K, ev in rows:
    d = (ev - pe) / (K - pk)
    print(f"iterations {pk + 1:2d}-{K:2d}: per pair-iteration  algorithmic {d[1]:9.0f}  visited {d[3]:9.0f}  fully evaluated {d[4]:9.0f}  geometry-only {d[5]:9.0f}")
    pk, pe = K, ev
