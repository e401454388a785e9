"""Sweep of the batch slicing (concurrent slices on separate streams): device-resident and host-buffer throughput."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from icp_b200 import algorithms as alg, capi, synth

L = capi.lib()
ctx = capi.Context(0)
base = ctx.upload(synth.base_landmarks())
for n_pairs in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "256").split(",")]:
    b = alg.ICPBatch(ctx, n_pairs, 16384, 256)
    b.synthesize(base, 5000); ctx.sync()
    pb = 16384 * 32
    hF = capi.PinnedArray((n_pairs, 16384, 8), np.float32); hM = capi.PinnedArray((n_pairs, 16384, 8), np.float32)
    capi.check(L.icp_memcpy_d2h(ctx.h, hF.ptr, L.icp_batch_F(b.h), n_pairs * pb, 1))
    capi.check(L.icp_memcpy_d2h(ctx.h, hM.ptr, L.icp_batch_M(b.h), n_pairs * pb, 1))
    want = None
    for ns in (1, 2, 3, 4, 6, 8):
        b.set_slices(ns)
        for _ in range(2): b.register(40)
        ctx.sync(); ctx.timer_start()
        for _ in range(4): b.register(40)
        ms = ctx.timer_stop() / 4
        T = b.read_poses()
        if want is None: want = T
        ok = np.array_equal(T.view(np.uint32), want.view(np.uint32))
        for _ in range(2): b.register_host(hF.ptr, hM.ptr, 40, ns)
        ctx.sync(); t0 = time.perf_counter()
        for _ in range(4): T2 = b.register_host(hF.ptr, hM.ptr, 40, ns)
        e2e = (time.perf_counter() - t0) / 4
        ok2 = np.array_equal(T2.view(np.uint32), want.view(np.uint32))
        print(f"pairs {n_pairs} slices {ns}: resident {ms:.2f} ms = {n_pairs/ms*1e3:.0f} pairs/s ({ok}); host {e2e*1e3:.2f} ms = {n_pairs/e2e:.0f} pairs/s ({ok2})", flush=True)
    b.close()
