"""Per-kernel times and kernel-D phase clocks of ONE large registration (default 307200 points / 1024 representatives)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth
m, nr, W, H = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (307200, 1024, 640, 480)))
ctx = capi.Context(0)
F = synth.grid_cloud(W, H)
F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
s.init(m, nr, 2e2, 1e-6, W, H)
s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
s.buildRBC(); ctx.sync()
ctx.timer_start(); s.run(20, variant=1); t = ctx.timer_stop() * 1e3 / 20
prof = s.debug("prof", np.uint64, 64)
clk = [int(x) for x in prof[:7]]
g = lambda k, i: int(prof[16 + 8 * k + i])
print(json.dumps({"m": m, "nr": nr, "us_per_iter": round(t, 2),
                  "D_phase_cycles": {"sum_w": clk[1] - clk[0], "means": clk[2] - clk[1], "S_partials": clk[3] - clk[2], "S_finish": clk[4] - clk[3],
                                     "power_method": clk[6] - clk[4], "accumulate": clk[5] - clk[6]},
                  "timeline_ns": {"A_end": g(0, 6) - g(0, 0), "B_start": g(1, 0) - g(0, 0), "C_start": g(2, 0) - g(0, 0), "D_start": g(3, 0) - g(0, 0), "D_end": g(3, 6) - g(0, 0),
                                  "A_end_last_CTA": int(prof[48]) - g(0, 0), "C_end_last_CTA": int(prof[50]) - g(0, 0)}}))
