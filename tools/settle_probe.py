"""How many stage-2 evaluations the settle test leaves per iteration (batch engine, ICP_B200_BATCH_EVALS=1)."""
import os, sys
os.environ["ICP_B200_BATCH_EVALS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth
ctx = capi.Context(0)
b = alg.ICPBatch(ctx, 12, 16384, 256)
base = ctx.upload(synth.base_landmarks())
b.synthesize(base, 5000)
prev = np.zeros(4, np.uint64)
for K in list(range(1, 13)) + [16, 20, 30, 40]:
    # evals accumulate inside one registration; a registration restarts from the build
    capi.check(capi.lib().icp_memset(ctx.h, capi.lib().icp_batch_debug_ptr(b.h, b"evals@3"), 0, 32))
    b.register(K); ctx.sync()
    ev = b.debug("evals", np.uint64, 4, pair=3)
    nnd = b.debug("nnd", np.float32, 16384, pair=3)
    print(K, "evals [e1, e2 algorithmic, e1 executed, e2 executed] cumulative:", ev, " bounds > 0:", int((nnd > 0).sum()), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez(os.path.join(ROOT, "gpurun_out", "pair3.npz"), F=b.debug("F", np.float32, (16384, 8), pair=3), M=b.debug("M", np.float32, (16384, 8), pair=3))
