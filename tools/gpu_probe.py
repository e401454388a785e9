"""First-contact probe on a B200: micro-benchmarks + per-variant iteration latency.  Not part of the test-suite."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from icp_b200 import algorithms as alg, capi, synth

ctx = capi.Context(0)
L = capi.lib()
out = {"sm_count": ctx.sm_count, "cc": ctx.cc, "clock_khz": ctx.clock_khz, "l2": ctx.l2_bytes}
r = (C.c_double * 4)()
capi.check(L.icp_measure_fp32_rates(ctx.h, r))
out["fp32_rates_tflops"] = dict(scalar_mul_add=r[0] / 1e12, packed_mul2_add2=r[1] / 1e12, scalar_ffma=r[2] / 1e12, packed_ffma2=r[3] / 1e12)
a, b = C.c_float(), C.c_float()
capi.check(L.icp_measure_launch_floor(ctx.h, C.byref(a), C.byref(b)))
out["launch_floor_us"] = dict(stream=a.value, graph_node=b.value)
print(json.dumps(out)); sys.stdout.flush()

F, M, _, _ = synth.known_transform_pair(seed=42)
res = {}
for mode, mname in ((0, "staged"), (1, "fused")):
    s = alg.ICPStep(ctx, 1, 1); s.init(16384, 256, 2e2, 1e-6); s.set_mode(mode)
    s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
    for variant, vname in ((0, "stream"), (1, "graph"), (2, "while")):
        try:
            ts = []
            for rep in range(6):
                s.reset(); s.buildRBC(); ctx.sync()
                ctx.timer_start(); s.run(40, variant=variant); ms = ctx.timer_stop()
                ts.append(ms * 1e3 / 40)
            res[f"{mname}_{vname}_us_per_iter"] = [round(t, 2) for t in ts]
        except Exception as e:
            res[f"{mname}_{vname}"] = "ERR " + str(e)
    if mode == 0:
        s.reset(); s.buildRBC(); s.run(3); ctx.sync()
        res["staged_stage_ms"] = s.run_timed()
    if mode == 1:
        s.reset(); s.buildRBC(); s.run(2, variant=0); ctx.sync()
        pr = s.debug("prof", np.uint64, 8).astype(np.int64)
        res["fused_D_phase_cycles"] = dict(sumw=int(pr[1] - pr[0]), means=int(pr[2] - pr[1]), sij=int(pr[3] - pr[2]),
                                           level2=int(pr[4] - pr[3]), solve=int(pr[5] - pr[4]), pm_iters=int(pr[7]))
    s.set_count_evals(True); s.reset(); s.buildRBC(); s.run(40); ctx.sync()
    res[f"{mname}_evals"] = s.eval_counts()
    s.close()
print(json.dumps(res)); sys.stdout.flush()

# SVD path and batch throughput
s = alg.ICPStep(ctx, 0, 1); s.init(16384, 256, 2e2, 1e-6); s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
ts = []
for rep in range(4):
    s.reset(); s.buildRBC(); ctx.sync(); ctx.timer_start(); s.run(40); ts.append(ctx.timer_stop() * 1e3 / 40)
print(json.dumps({"fused_svd_us_per_iter": ts})); s.close()
base = ctx.upload(synth.base_landmarks())
for n_pairs in (16, 64, 256):
    b = alg.ICPBatch(ctx, n_pairs, 16384, 256)
    b.synthesize(base, 5000); ctx.sync()
    ts = []
    for rep in range(3):
        ctx.timer_start(); b.register(40); ts.append(ctx.timer_stop())
    print(json.dumps({"batch_pairs": n_pairs, "ms": ts, "pairs_per_s": n_pairs / (min(ts) * 1e-3), "us_per_pair_iter": min(ts) * 1e3 / n_pairs / 40}))
    sys.stdout.flush()
    b.close()
