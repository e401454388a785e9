"""Generates tests/golden/stage_vectors.npz: small seeded inputs and the outputs of the REFERENCE's own CPU helpers
(/root/reference/include/ICP/tests/helper_funcs.hpp compiled in place into oracle/_ref/libicp_ref.so, see
oracle/Makefile).  Run in the build container, where /root/reference exists; the .npz is committed so that the oracle
stays pinned on machines without the reference tree (the GPU box)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import pyoracle as po  # noqa: E402

R = po.ref_lib()
assert R is not None, "oracle/_ref/libicp_ref.so missing: run `make -C oracle ref` where /root/reference exists"
rng = np.random.default_rng(20261017)
n = 2048
out = {}

cloud = rng.uniform(0, 1, (480 * 640, 8)).astype(np.float32)
lms = np.empty((16384, 8), np.float32); R.ref_ICPLMs(cloud.reshape(-1), lms.reshape(-1))
# keep the golden small: store the cloud seed, a checksum of the input and the sampled output rows 0, 5000, 16383
out["lms_seed"] = np.array([20261017]); out["lms_rows"] = lms[[0, 5000, 16383]].copy()
out["lms_sum"] = np.array([lms.astype(np.float64).sum()])
for nr in (256, 512, 1024):
    reps = np.empty((nr, 8), np.float32); R.ref_ICPReps(lms.reshape(-1), reps.reshape(-1), nr)
    out[f"reps_{nr}"] = reps[:: nr // 16].copy()

M = rng.uniform(0, 255, (n, 8)).astype(np.float32)
Tq = np.array([0.5144, 0.5743, 0.5632, 0.2973, 12.5, 200.25, 31.0, 0.73], np.float32)
o = np.empty_like(M); R.ref_ICPTransformQ(M.reshape(-1).copy(), o.reshape(-1), Tq.copy(), n)
out["tq_in"], out["tq_T"], out["tq_out"] = M, Tq, o.copy()
s = 0.61
Tm = np.array([s * .871238, s * -.276687, s * .405449, 3.0, s * .405449, s * .871238, s * -.276687, 77.0,
               s * -.276687, s * .405449, s * .871238, 140.0, 0, 0, 0, 1], np.float32)
o2 = np.empty_like(M); R.ref_ICPTransformM(M.reshape(-1).copy(), o2.reshape(-1), Tm.copy(), n)
out["tm_T"], out["tm_out"] = Tm, o2.copy()

dist = rng.uniform(1e-6, 255e-6, n).astype(np.float32)
D = np.zeros(n, dtype=[("d", np.float32), ("i", np.uint32)]); D["d"] = dist
W = np.empty(n, np.float32); sw = C.c_double(); R.ref_ICPWeights(D.ctypes.data, W, C.byref(sw), n)
out["w_dist"], out["w_out"] = dist, W.copy()

F = rng.uniform(0, 1, (n, 8)).astype(np.float32); Mm = rng.uniform(0, 1, (n, 8)).astype(np.float32)
mean = np.empty(8, np.float32); R.ref_ICPMean(F.reshape(-1), Mm.reshape(-1), mean, n)
out["mean_F"], out["mean_M"], out["mean_out"] = F, Mm, mean.copy()
Wt = rng.uniform(0, 1, n).astype(np.float32)
wmean = np.empty(8, np.float32); R.ref_ICPMeanWeighted(F.reshape(-1), Mm.reshape(-1), wmean, Wt, n)
out["wmean_W"], out["wmean_out"] = Wt, wmean.copy()
DF = np.empty((n, 4), np.float32); DM = np.empty((n, 4), np.float32)
R.ref_ICPDevs(F.reshape(-1), Mm.reshape(-1), DF.reshape(-1), DM.reshape(-1), wmean, n)
out["devs_DF"], out["devs_DM"] = DF.copy(), DM.copy()
DMx = rng.uniform(-1000, 1000, (n, 4)).astype(np.float32); DFx = rng.uniform(-1000, 1000, (n, 4)).astype(np.float32)
S = np.empty(11, np.float32); R.ref_ICPS(DMx.reshape(-1), DFx.reshape(-1), S, n, 1e-6)
Sw = np.empty(11, np.float32); R.ref_ICPSw(DMx.reshape(-1), DFx.reshape(-1), Wt, Sw, n, 1e-6)
out["s_DM"], out["s_DF"], out["s_out"], out["sw_out"] = DMx, DFx, S.copy(), Sw.copy()

# power method: the reference KAT + 6 random S matrices through the reference helper
KAT_S = np.array([0.00168053, 0.000131408, -0.000775179, 0.000156595, 0.00102674, -0.000563479,
                  -0.000722137, -0.000559463, 0.00246661, 0.00521271, 0.00515292], np.float32)
KAT_M = np.array([-33.9694, -17.6421, 1494.22, 0., -44.8322, -19.3835, 1485.93, 0.], np.float32)
Ss, Ts = [KAT_S], []
for i in range(6):
    A = rng.normal(size=(3, 3)); S3 = (A @ A.T) * 1e-3 + rng.normal(size=(3, 3)) * 1e-4
    if i % 3 == 2:
        S3 = -S3
    Ss.append(np.r_[S3.reshape(-1), 0.0052, 0.0051].astype(np.float32))
for Sv in Ss:
    Tk = np.zeros(8, np.float32); R.ref_ICPPowerMethod(Sv.copy(), KAT_M.copy(), Tk); Ts.append(Tk)
out["pm_S"], out["pm_means"], out["pm_Tk"] = np.stack(Ss), KAT_M, np.stack(Ts)

a = rng.uniform(0, 1, (5, 1024)).astype(np.float32)
rs = np.empty(5, np.float32); R.ref_ReduceSum(a.reshape(-1), rs, 1024, 5)
rmin = np.empty(5, np.float32); R.ref_ReduceMin(a.reshape(-1), rmin, 1024, 5)
out["red_in"], out["red_sum"], out["red_min"] = a, rs.copy(), rmin.copy()
ai = rng.integers(0, 10000, (3, 777)).astype(np.int32)
inc = np.empty_like(ai); exc = np.empty_like(ai)
R.ref_InScan(ai.reshape(-1).copy(), inc.reshape(-1), 777, 3); R.ref_ExScan(ai.reshape(-1).copy(), exc.reshape(-1), 777, 3)
out["scan_in"], out["scan_inc"], out["scan_exc"] = ai, inc.copy(), exc.copy()

np.savez_compressed(os.path.join(HERE, "stage_vectors.npz"), **out)
print("wrote", os.path.join(HERE, "stage_vectors.npz"), {k: v.shape for k, v in out.items()})
