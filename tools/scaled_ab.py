"""Single-registration engine at the BASELINE sizes (16384/256 and the scaled sets of configs[3]): device-timed us per ICP
iteration over 40 iterations (best of 3), executed / algorithmic evaluation counts.  One JSON line per size.
usage: [ICP_B200_SETTLE=0|1 ...] python tools/scaled_ab.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
ITERS = 40
ctx = capi.Context(0)
for m, nr, lm in [(16384, 256, (128, 128)), (65536, 512, (256, 256)), (65536, 1024, (256, 256)), (307200, 512, (640, 480)), (307200, 1024, (640, 480))]:
    F = synth.base_landmarks() if m == 16384 else synth.grid_cloud(*lm)
    F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
    s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
    s.init(m, nr, 2e2, 1e-6, lm[0], lm[1])
    s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
    ts = []
    for rep in range(4):
        s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(ITERS); ts.append(ctx.timer_stop() * 1e3 / ITERS)
    s.set_count_evals(True)
    s.reset(); s.buildRBC(); s.run(ITERS); ctx.sync()
    e1, e2 = s.eval_counts(); e1x, e2x = s.stage1_executed(), s.stage2_executed()
    T = s.debug("T", np.float32, 8)
    s.close()
    print(json.dumps({"m": m, "nr": nr, "us_per_iter": round(min(ts[1:]), 2), "e1x_frac": round(e1x / e1, 4), "e2x_frac": round(e2x / e2, 4),
                      "T_bits": int(T.view(np.uint32).sum())}), flush=True)
