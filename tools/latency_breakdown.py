"""Latency-mode breakdown for ONE pair (16384/256): per-kernel device times, engine variants, phases inside kernel D."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth

ctx = capi.Context(0)
out = {}
base = ctx.upload(synth.base_landmarks())
for rot, name in ((capi.ROT_POWER_METHOD, "power_method"), (capi.ROT_EIGEN, "svd")):
    b = alg.ICPBatch(ctx, 1, 16384, 256, rot=rot)
    b.synthesize(base, 5000)
    b.register(3); ctx.sync()
    out[name] = {"cfg": b.config(), "kernel_us": {k: round(1e3 * b.time_kernel(i, 20), 2) for i, k in enumerate("ABCD")}}
    b.close()
F, M, _, _ = synth.known_transform_pair(seed=42)
for rot, name in ((capi.ROT_POWER_METHOD, "power_method"), (capi.ROT_EIGEN, "svd")):
    s = alg.ICPStep(ctx, rot, 1); s.init(16384, 256, 2e2, 1e-6)
    s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
    for variant, vn in ((0, "stream"), (1, "graph_unrolled"), (2, "graph_while")):
        ts = []
        for rep in range(4):
            s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(40, variant=variant); ts.append(ctx.timer_stop() * 1e3 / 40)
        out[name][f"us_per_iter_{vn}"] = round(min(ts), 2)
    prof = s.debug("prof", np.uint64, 64)
    clk = [int(x) for x in prof[:6]]
    out[name]["D_phase_cycles(last iteration)"] = {"sum_w": clk[1] - clk[0], "means": clk[2] - clk[1], "S_partials": clk[3] - clk[2],
                                                  "S_finish": clk[4] - clk[3], "solve+accumulate": clk[5] - clk[4], "pm_iterations": int(prof[7]),
                                                  "power_method_only": (int(prof[6]) - clk[4]) if int(prof[7]) else 0}
    g = lambda k, i: int(prof[16 + 8 * k + i])
    out[name]["timeline_ns(last iteration, CTA 0)"] = {
        "A_start": 0, "A_end": g(0, 6) - g(0, 0), "B_start": g(1, 0) - g(0, 0), "C_start": g(2, 0) - g(0, 0), "C_end": g(2, 6) - g(0, 0),
        "D_start": g(3, 0) - g(0, 0), "D_end": g(3, 6) - g(0, 0),
        "A_end_last_CTA": int(prof[48]) - g(0, 0), "C_end_last_CTA": int(prof[50]) - g(0, 0)}
    out[name]["A_phase_cycles(CTA 0)"] = {"stage_reps": g(0, 2) - g(0, 1), "pruned_pass": g(0, 3) - g(0, 2), "full_scan_pass": g(0, 4) - g(0, 3), "rank_store": g(0, 5) - g(0, 4)}
    out[name]["C_phase_cycles(CTA 0)"] = {"scan_Nq": g(2, 2) - g(2, 1), "pass1": g(2, 3) - g(2, 2), "items+pass2": g(2, 4) - g(2, 3), "item_loop(thread 0)": g(2, 5) - g(2, 4)}
    s.close()
print(json.dumps(out, indent=1))
