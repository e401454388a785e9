"""Latency mode, one pair (16384 / 256, power method + weighted): device-timed us per ICP iteration of every loop engine --
0 plain stream launches, 1 unrolled CUDA graph, 2 conditional WHILE graph, 3 persistent cooperative kernel -- 40 iterations,
best of 6 repetitions each (buildRBC before every repetition, L2 warm).  Prints one JSON line."""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth

ITERS = 40
ctx = capi.Context(0)
F, M, _, _ = synth.batch_pair(9000)
out = {}
for rot_name, rot in (("power_method", capi.ROT_POWER_METHOD), ("svd", capi.ROT_EIGEN)):
    res = {}
    for name, variant, env in (("stream", 0, {}), ("graph_unrolled", 1, {}), ("graph_while", 2, {}),
                               ("persistent_512", 3, {"ICP_B200_PERSIST_T": "512"}), ("persistent_512_flat_barrier", 3, {"ICP_B200_PERSIST_T": "512", "ICP_B200_PERSIST_HIER": "0"}),
                               ("persistent_1024", 3, {"ICP_B200_PERSIST_T": "1024"}),
                               ("persistent_512_64ctas", 3, {"ICP_B200_PERSIST_T": "512", "ICP_B200_PERSIST_CTAS": "64"})):
        for k, v in env.items():
            os.environ[k] = v
        s = alg.ICPStep(ctx, rot, capi.W_WEIGHTED)
        s.init(16384, 256, 2e2, 1e-6)
        s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, M)
        ts = []
        try:
            for rep in range(6):
                s.reset(); s.buildRBC(); ctx.sync()
                ctx.timer_start(); s.run(ITERS, variant=variant); ts.append(ctx.timer_stop() * 1e3 / ITERS)
            res[name] = round(min(ts[1:]), 2)
            res[name + "_T"] = s.debug("T", np.float32, 8).view(np.uint32).tolist()
        except Exception as e:
            res[name] = f"unavailable: {e}"
        s.close()
        for k in env:
            os.environ.pop(k, None)
    poses = {tuple(v) for k, v in res.items() if k.endswith("_T")}
    res = {k: v for k, v in res.items() if not k.endswith("_T")}
    res["all_engines_bit_identical"] = len(poses) == 1
    out[rot_name] = res
print(json.dumps(out))
