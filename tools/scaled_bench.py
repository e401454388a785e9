"""BASELINE config 4: scaled landmark sets (|F|=|M|=65536 and 307200, |R|=512/1024) on one GPU, one registration at a
time: device-timed us per ICP iteration, executed / algorithmic distance evaluations, fraction of the FP32 (non-fused
mul/add) and HBM rooflines.  Also the canonical 16384/256 case for comparison.  Prints one JSON line per configuration."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth

ITERS = 20
ctx = capi.Context(0)
L = capi.lib()
rates = (C.c_double * 4)()
capi.check(L.icp_measure_fp32_rates(ctx.h, rates))
fp32_peak = rates[0]
for m, nr, lm in [(16384, 256, (128, 128)), (65536, 512, (256, 256)), (65536, 1024, (256, 256)), (307200, 512, (640, 480)), (307200, 1024, (640, 480))]:
    F = synth.base_landmarks() if m == 16384 else synth.grid_cloud(*lm)
    F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
    s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
    s.init(m, nr, 2e2, 1e-6, lm[0], lm[1])
    s.write(capi.MEM_D_IN_F, F2); s.write(capi.MEM_D_IN_M, M_)
    ts = []
    for rep in range(4):
        s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(ITERS); ts.append(ctx.timer_stop() * 1e3 / ITERS)
    us = float(np.median(ts[1:]))
    s.set_count_evals(True)
    s.reset(); s.buildRBC(); s.run(ITERS); ctx.sync()
    e1, e2 = s.eval_counts()
    e1x, e2x = s.stage1_executed(), s.stage2_executed()
    s.close()
    flop_alg = (25.0 * (e1 + e2) / ITERS + 75.0 * m)
    flop_exec = (25.0 * (e1x + e2x) / ITERS + 75.0 * m)
    bytes_alg = 32.0 * m * 2 + 40.0 * nr + 64
    print(json.dumps({"m": m, "nr": nr, "us_per_icp_iteration": round(us, 2), "stage1_evals_per_iter": e1 // ITERS, "stage1_executed_per_iter": e1x // ITERS,
                      "stage2_evals_per_iter": e2 // ITERS, "stage2_executed_per_iter": e2x // ITERS,
                      "fp32_roofline_frac_algorithmic": round(flop_alg / (us * 1e-6) / fp32_peak, 4),
                      "fp32_roofline_frac_executed": round(flop_exec / (us * 1e-6) / fp32_peak, 4),
                      "fp32_peak_tflops": round(fp32_peak / 1e12, 2),
                      "hbm_gbs_algorithmic": round(bytes_alg / (us * 1e-6) / 1e9, 1), "hbm_frac_of_6650": round(bytes_alg / (us * 1e-6) / 1e9 / 6650.0, 4)}), flush=True)
