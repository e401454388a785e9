"""CPU tests that PIN the oracle (oracle/icp_oracle.cpp) to the reference:
  * the reference's known-answer vector for the power method (tests/testsICP.cpp:1008-1046),
  * committed golden vectors produced by the reference's own CPU helpers (tests/golden/make_golden.py),
  * the live helpers in oracle/_ref when the reference tree is present (build container),
plus property tests of the restated RBC (brute-force agreement inside the chosen list, stable order, partition)."""
import os

import numpy as np
import pytest

from util import assert_bits_equal, rng_points

EPS = float(np.finfo(np.float32).eps)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage_vectors.npz"))

KAT_S = np.array([0.00168053, 0.000131408, -0.000775179, 0.000156595, 0.00102674, -0.000563479,
                  -0.000722137, -0.000559463, 0.00246661, 0.00521271, 0.00515292], np.float32)
KAT_MEANS = np.array([-33.9694, -17.6421, 1494.22, 0., -44.8322, -19.3835, 1485.93, 0.], np.float32)
KAT_SVD_TK = np.array([0.00111412, 0.00730956, -0.00647493, 0.999952, -10.4598, 4.74009, -0.762817, 1.00578], np.float32)


def test_power_method_known_answer(po):
    Tk, iters = po.power_method(KAT_S, KAT_MEANS)
    assert np.abs(Tk - KAT_SVD_TK).max() < 42000 * EPS          # testsICP.cpp:1050
    assert 20 < iters < 200
    # scale follows the KERNEL convention S9/S10 = |f|^2/|m|^2
    assert abs(Tk[7] - np.sqrt(KAT_S[9] / KAT_S[10])) < 1e-6


def test_svd_solve_known_answer(po):
    Tk, Rk = po.svd_solve(KAT_S, KAT_MEANS)
    assert np.abs(Tk - KAT_SVD_TK).max() < 42000 * EPS
    U, s, Vt = np.linalg.svd(KAT_S[:9].astype(np.float64).reshape(3, 3))
    assert np.abs(Rk - Vt.T @ U.T).max() < 2e-6                   # R = V U^T is the polar factor: unique
    assert abs(np.linalg.det(Rk.astype(np.float64)) - 1) < 1e-5


def test_svd_reflection_fix(po):
    rng = np.random.default_rng(3)
    for _ in range(50):
        S3 = rng.normal(size=(3, 3)) * 1e-3
        S = np.r_[S3.reshape(-1), 0.005, 0.005].astype(np.float32)
        Tk, Rk = po.svd_solve(S, KAT_MEANS)
        R64 = Rk.astype(np.float64)
        assert np.abs(R64 @ R64.T - np.eye(3)).max() < 5e-6
        assert np.linalg.det(R64) > 0.999                           # det fix of algorithms.cpp:3889-3894


def test_golden_sampling(po):
    rng = np.random.default_rng(int(GOLD["lms_seed"][0]))
    cloud = rng.uniform(0, 1, (480 * 640, 8)).astype(np.float32)
    lms = po.get_lms(cloud)
    assert_bits_equal(lms[[0, 5000, 16383]], GOLD["lms_rows"], "lms rows")
    assert lms.astype(np.float64).sum() == GOLD["lms_sum"][0]
    for nr in (256, 512, 1024):
        assert_bits_equal(po.get_reps(lms, 128, 128, nr)[:: nr // 16], GOLD[f"reps_{nr}"], f"reps {nr}")


def test_golden_transforms(po):
    assert_bits_equal(po.transform_q(GOLD["tq_in"], GOLD["tq_T"]), GOLD["tq_out"], "transform Q")
    assert_bits_equal(po.transform_m(GOLD["tq_in"], GOLD["tm_T"]), GOLD["tm_out"], "transform M")


def test_golden_weights_means_devs_S(po):
    n = len(GOLD["w_dist"])
    W, sw = po.weights(GOLD["w_dist"])
    assert_bits_equal(W, GOLD["w_out"], "weights")
    assert abs(sw - GOLD["w_out"].astype(np.float64).sum()) < 4200 * EPS            # testsICP.cpp:282-286
    F, M = GOLD["mean_F"], GOLD["mean_M"]
    assert np.abs(po.mean(F, M) - GOLD["mean_out"]).max() < 420000 * EPS            # :369  (tree vs serial order)
    Wt = GOLD["wmean_W"]
    wm = po.mean_weighted(F, M, Wt, float(Wt.astype(np.float64).sum()))
    assert np.abs(wm - GOLD["wmean_out"]).max() < 420000 * EPS                      # :461
    DF, DM = po.devs(F, M, GOLD["wmean_out"])
    assert_bits_equal(DF, GOLD["devs_DF"], "DF"); assert_bits_equal(DM, GOLD["devs_DM"], "DM")
    # S: rows 0..8 agree with the helper; rows 9/10 are swapped between helper and kernel (kernel = truth)
    for key, w in (("s_out", None), ("sw_out", Wt)):
        S = po.sij(GOLD["s_DM"], GOLD["s_DF"], w, 1e-6)
        assert np.abs(S[:9] - GOLD[key][:9]).max() < 4200 * EPS                      # :653, :752
        assert abs(S[9] - GOLD[key][10]) < 4200 * EPS and abs(S[10] - GOLD[key][9]) < 4200 * EPS
    assert n == 2048


def test_golden_power_method_bit_exact(po):
    for S, Tk in zip(GOLD["pm_S"], GOLD["pm_Tk"]):
        got, _ = po.power_method(S, GOLD["pm_means"])
        assert_bits_equal(got, Tk, "power method vs reference helper")


def test_golden_reduce_scan(po):
    a = GOLD["red_in"]
    assert np.abs(po.reduce_sum_f(a) - GOLD["red_sum"]).max() < 42000 * EPS          # testsReduce.cpp:263
    assert_bits_equal(po.reduce_min_f(a), GOLD["red_min"], "reduce min")
    assert np.array_equal(po.scan_i(GOLD["scan_in"], True), GOLD["scan_inc"])
    assert np.array_equal(po.scan_i(GOLD["scan_in"], False), GOLD["scan_exc"])


def test_live_reference_helpers_when_present(po):
    """Build container only: the same comparisons against the helpers compiled from /root/reference right now."""
    R = po.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(99)
    M = rng.uniform(0, 255, (4096, 8)).astype(np.float32)
    T = np.array([0.5144, 0.5743, 0.5632, 0.2973, 1.0, 2.0, 3.0, 0.5], np.float32)
    out = np.empty_like(M); R.ref_ICPTransformQ(M.reshape(-1).copy(), out.reshape(-1), T.copy(), len(M))
    assert_bits_equal(po.transform_q(M, T), out, "live transform Q")
    Tk = np.zeros(8, np.float32); R.ref_ICPPowerMethod(KAT_S.copy(), KAT_MEANS.copy(), Tk)
    assert_bits_equal(po.power_method(KAT_S, KAT_MEANS)[0], Tk, "live power method")


# ---------------------------------------------------------------------------------------------------------
# RBC restatement: properties (parity with the real RandomBallCover library is unpinned)
# ---------------------------------------------------------------------------------------------------------
def test_rbc_construct_properties(po):
    rng = np.random.default_rng(5)
    n, nr, a = 3000, 48, 2e2
    X = rng_points(rng, n)
    X[100] = X[7]; X[2000] = X[7]
    R = X[rng.choice(n, nr, replace=False)]
    rb = po.rbc_construct(X, R, a)
    fg, fp = po.metric_weights(a)
    assert abs(fg - 1 / 201) < 1e-9 and abs(fp - 200 / 201) < 1e-7
    # brute force argmin with numpy in the oracle's evaluation order
    d = X[:, None, :] - R[None, :, :]
    sq = d * d
    g = ((sq[..., 0] + sq[..., 1]) + sq[..., 2]) + sq[..., 3]
    p = ((sq[..., 4] + sq[..., 5]) + sq[..., 6]) + sq[..., 7]
    D = np.float32(fg) * g + np.float32(fp) * p
    assert np.array_equal(rb["rep_id"], D.argmin(1).astype(np.uint32))       # argmin => first (lowest) index on ties
    assert rb["N"].sum() == n and np.array_equal(rb["O"], np.r_[0, np.cumsum(rb["N"])[:-1]].astype(np.uint32))
    assert sorted(rb["perm"].tolist()) == list(range(n))                       # a permutation
    for r in range(nr):                                                         # stable: ascending original index per list
        seg = rb["perm"][rb["O"][r]: rb["O"][r] + rb["N"][r]]
        assert np.all(np.diff(seg.astype(np.int64)) > 0) and np.all(rb["rep_id"][seg] == r)
    assert_bits_equal(rb["Xp"], X[rb["perm"]], "Xp")


def test_rbc_search_is_exact_inside_the_list(po):
    rng = np.random.default_rng(6)
    n, nr, a = 2048, 32, 1.0
    X = rng_points(rng, n); Q = rng_points(rng, n)
    R = X[:: n // nr][:nr].copy()
    rb = po.rbc_construct(X, R, a)
    sr = po.rbc_search(Q, R, a, rb["Xp"], rb["O"], rb["N"])
    fg, fp = po.metric_weights(a)
    assert sorted(sr["qperm"].tolist()) == list(range(n))
    for p in range(0, n, 37):
        q = sr["Qp"][p]; r = sr["q_rep"][sr["qperm"][p]]
        lst = rb["Xp"][rb["O"][r]: rb["O"][r] + rb["N"][r]]
        dd = q[None, :] - lst
        sq = dd * dd
        D = np.float32(fg) * (((sq[:, 0] + sq[:, 1]) + sq[:, 2]) + sq[:, 3]) + np.float32(fp) * (((sq[:, 4] + sq[:, 5]) + sq[:, 6]) + sq[:, 7])
        assert sr["nn_id"][p] == rb["O"][r] + int(D.argmin())
        assert sr["nn_dist"][p] == D.min()
        assert_bits_equal(sr["NN"][p], rb["Xp"][sr["nn_id"][p]], "NN point")


def test_registration_recovers_a_small_motion(po):
    from icp_b200 import synth
    F, M, R_gt, t_gt = synth.known_transform_pair(seed=21, deg=1.0, t=(6.0, -4.0, 3.0), outliers=0.0, xyz_sigma=0.2)
    res = po.icp_register(F, M, 128, 128, 256, fixed_iters=40)
    T = res["T16"]
    assert np.linalg.norm(T[:3, 3] - t_gt) < 0.5 * np.linalg.norm(t_gt)
    assert abs(np.linalg.det(T[:3, :3].astype(np.float64)) - 1) < 5e-2


def test_driver_stops_like_icp_check(po):
    from icp_b200 import synth
    F, M, _, _ = synth.known_transform_pair(seed=12, deg=0.4, t=(2.0, -1.0, 1.5))
    r1 = po.icp_register(F, M, 128, 128, 256, fixed_iters=0, max_iterations=40, angle_thr=0.05, trans_thr=0.5, dumps=True)
    assert 1 <= r1["k"] <= 40
    if r1["k"] < 40:     # converged: the last increment is below both thresholds, the one before is not
        qk, tk = r1["Tk_hist"][-1][:4], r1["Tk_hist"][-1][4:7]
        ang = np.degrees(2 * np.arctan2(np.linalg.norm(qk[:3]), qk[3]))
        assert ang < 0.05 and np.linalg.norm(tk) < 0.5
    r2 = po.icp_register(F, M, 128, 128, 256, fixed_iters=0, max_iterations=3)
    assert r2["k"] == 3
