"""Kernel-config sweep (one process per config; knobs via env).  Prints per-kernel ms for a batch and single-pair us/iter."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os
sys.path.insert(0, %r)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
ctx = capi.Context(0)
base = ctx.upload(synth.base_landmarks())
n = int(os.environ.get("TUNE_PAIRS", "64"))
out = {}
if n > 0:
    b = alg.ICPBatch(ctx, n, 16384, 256)
    b.synthesize(base, 5000); b.register(3); ctx.sync()
    out["cfg"] = b.config()
    out["ms"] = [round(b.time_kernel(w, 40), 4) for w in range(4)]
    ctx.timer_start(); b.register(40); out["us_per_pair_iter"] = round(ctx.timer_stop() * 1e3 / n / 40, 3)
    b.close()
else:
    F, M, _, _ = synth.known_transform_pair(seed=42)
    s = alg.ICPStep(ctx, 1, 1); s.init(16384, 256, 2e2, 1e-6); s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
    ts = []
    for rep in range(4):
        s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(40); ts.append(round(ctx.timer_stop() * 1e3 / 40, 2))
    out["single_us_per_iter"] = ts
print(json.dumps(out))
''' % ROOT

def run(env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True)
    print(json.dumps(env), "->", r.stdout.strip() or r.stderr.strip()[-300:])
    sys.stdout.flush()

if __name__ == "__main__":
    for L in (1, 4, 8, 32):
        for QC in (256, 512):
            run({"TUNE_PAIRS": 64, "ICP_B200_L": L, "ICP_B200_QC": QC})
    for S in (2, 4, 8, 16):
        for QB in (128, 256, 512):
            run({"TUNE_PAIRS": 64, "ICP_B200_S": S, "ICP_B200_QB": QB})
    for L in (4, 8, 16):
        for QC in (32, 64, 128):
            run({"TUNE_PAIRS": 0, "ICP_B200_L": L, "ICP_B200_QC": QC})
    for S in (8, 16, 32):
        for QB in (56, 112, 224):
            run({"TUNE_PAIRS": 0, "ICP_B200_S": S, "ICP_B200_QB": QB})
