"""GPU parity tests of the individual pipeline stages, through the C ABI, against the CPU oracle.
They mirror the reference's own stage tests (/root/reference/tests/testsICP.cpp, testsReduce.cpp,
testsScan.cpp: same sizes and input distributions, fixed seeds) but demand bit-exact results instead of the
reference's epsilon bands, since oracle and kernels share one evaluation order."""
import numpy as np
import pytest

from util import assert_bits_equal, rng_points

pytestmark = pytest.mark.gpu

M = 16384
NR = 256


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


def test_getLMs(ctx, po, alg):                      # testsICP.cpp:66-141
    rng = np.random.default_rng(1)
    cloud = rng.uniform(0, 1, (640 * 480, 8)).astype(np.float32)
    lm = alg.ICPLMs(ctx); lm.init(); lm.write("D_IN", cloud); lm.run()
    assert_bits_equal(lm.read(), po.get_lms(cloud), "getLMs")


@pytest.mark.parametrize("nr,W,H", [(256, 128, 128), (512, 128, 128), (1024, 128, 128), (16, 128, 128),
                                    (512, 256, 256), (1024, 640, 480), (512, 640, 480)])
def test_getReps(ctx, po, alg, nr, W, H):           # testsICP.cpp:147-222 (+ generalised grids)
    rng = np.random.default_rng(2)
    lms = rng.uniform(0, 1, (W * H, 8)).astype(np.float32)
    r = alg.ICPReps(ctx); r.init(nr, W, H); r.write("D_IN", lms); r.run()
    assert_bits_equal(r.read(), po.get_reps(lms, W, H, nr), "getReps")


def test_getReps_config_errors(ctx, alg):
    r = alg.ICPReps(ctx); r.init(6)
    with pytest.raises(alg.ICPConfigError):
        r.run()                                      # not a multiple of 4 (algorithms.cpp:842)


@pytest.mark.parametrize("m", [M, 2, 1000, 65536])
def test_transform_quaternion(ctx, po, alg, m):      # testsICP.cpp:796-884
    rng = np.random.default_rng(3)
    pts = rng.uniform(0, 255, (m, 8)).astype(np.float32)
    T = np.array([0.5144, 0.5743, 0.5632, 0.2973, *rng.uniform(0, 255, 3), rng.uniform()], np.float32)
    t = alg.ICPTransform(ctx, "QUATERNION"); t.init(m); t.write("D_IN_M", pts); t.write("D_IN_T", T); t.run()
    assert_bits_equal(t.read(), po.transform_q(pts, T), "transformQ")


def test_transform_matrix(ctx, po, alg):             # testsICP.cpp:890-981
    rng = np.random.default_rng(4)
    pts = rng.uniform(0, 255, (M, 8)).astype(np.float32)
    s = 0.37
    T = np.array([s * .871238, s * -.276687, s * .405449, 17., s * .405449, s * .871238, s * -.276687, 201.,
                  s * -.276687, s * .405449, s * .871238, 99., 0, 0, 0, 1], np.float32)
    t = alg.ICPTransform(ctx, "MATRIX"); t.init(M); t.write("D_IN_M", pts); t.write("D_IN_T", T); t.run()
    assert_bits_equal(t.read(), po.transform_m(pts, T), "transformM")


@pytest.mark.parametrize("n", [M, 128, 2, 130, 4096, 65536, 307200])
def test_weights(ctx, po, alg, n):                   # testsICP.cpp:228-321
    rng = np.random.default_rng(5)
    d = np.zeros(n, alg.DIST_ID)
    d["dist"] = rng.uniform(1e-6, 255e-6, n).astype(np.float32) if n != 4096 else rng.uniform(0, 5000, n).astype(np.float32)
    d["id"] = rng.integers(0, n, n)
    w = alg.ICPWeights(ctx); w.init(n); w.write("D_IN", d); w.run()
    W, sw = w.read()
    Wo, swo = po.weights(d["dist"])
    assert_bits_equal(W, Wo, "W")
    assert np.float64(sw).view(np.uint64) == np.float64(swo).view(np.uint64), (sw, swo)
    assert abs(sw - W.astype(np.float64).sum()) < 4200 * np.finfo(np.float32).eps * max(1.0, n / 16384)   # reference band


def test_weights_config_errors(ctx, alg):
    w = alg.ICPWeights(ctx); w.init(3)
    with pytest.raises(alg.ICPConfigError):
        w.run()


@pytest.mark.parametrize("n", [M, 128, 2, 1000, 65536, 307200])
def test_mean(ctx, po, alg, n):                      # testsICP.cpp:327-406
    rng = np.random.default_rng(6)
    F, Mm = rng_points(rng, n), rng_points(rng, n)
    mn = alg.ICPMean(ctx, "REGULAR"); mn.init(n); mn.write("D_IN_F", F); mn.write("D_IN_M", Mm); mn.run()
    got = mn.read()
    assert_bits_equal(got, po.mean(F, Mm), "mean")
    ref = np.r_[F[:, :3].astype(np.float64).mean(0), 0, Mm[:, :3].astype(np.float64).mean(0), 0]
    assert np.abs(got - ref).max() < 420000 * np.finfo(np.float32).eps


@pytest.mark.parametrize("n", [M, 128, 1000, 65536])
def test_mean_weighted(ctx, po, alg, n):             # testsICP.cpp:412-498
    rng = np.random.default_rng(7)
    F, Mm = rng_points(rng, n, 0, 3000), rng_points(rng, n, 0, 3000)
    W = rng.uniform(0, 1, n).astype(np.float32)
    sw = np.array([float(W.astype(np.float64).sum())])
    mn = alg.ICPMean(ctx, "WEIGHTED"); mn.init(n)
    mn.write("D_IN_F", F); mn.write("D_IN_M", Mm); mn.write("D_IN_W", W); mn.write("D_IN_SUM_W", sw); mn.run()
    assert_bits_equal(mn.read(), po.mean_weighted(F, Mm, W, sw[0]), "mean_weighted")


def test_devs(ctx, po, alg):                         # testsICP.cpp:504-595
    rng = np.random.default_rng(8)
    F, Mm = rng_points(rng, M), rng_points(rng, M)
    mean = rng.uniform(0, 1, 8).astype(np.float32); mean[3] = 0; mean[7] = 0
    d = alg.ICPDevs(ctx); d.init(M); d.write("D_IN_F", F); d.write("D_IN_M", Mm); d.write("D_IN_MEAN", mean); d.run()
    DF, DM = d.read()
    DFo, DMo = po.devs(F, Mm, mean)
    assert_bits_equal(DF, DFo, "DF"); assert_bits_equal(DM, DMo, "DM")


@pytest.mark.parametrize("m,weighted", [(M, False), (M, True), (2048, True), (100, False), (65536, True), (307200, True), (1001 * 2, True)])
def test_sij(ctx, po, alg, m, weighted):             # testsICP.cpp:601-789
    rng = np.random.default_rng(9)
    DM = rng.uniform(-1000, 1000, (m, 4)).astype(np.float32)
    DF = rng.uniform(-1000, 1000, (m, 4)).astype(np.float32)
    W = rng.uniform(0, 1, m).astype(np.float32)
    s = alg.ICPS(ctx, "WEIGHTED" if weighted else "REGULAR"); s.init(m, 1e-6)
    s.write("D_IN_DEV_M", DM); s.write("D_IN_DEV_F", DF)
    if weighted:
        s.write("D_IN_W", W)
    s.run()
    assert_bits_equal(s.read(), po.sij(DM, DF, W if weighted else None, 1e-6), "Sij")


# the reference's known-answer vector (testsICP.cpp:1008-1046)
KAT_S = np.array([0.00168053, 0.000131408, -0.000775179, 0.000156595, 0.00102674, -0.000563479,
                  -0.000722137, -0.000559463, 0.00246661, 0.00521271, 0.00515292], np.float32)
KAT_MEANS = np.array([-33.9694, -17.6421, 1494.22, 0., -44.8322, -19.3835, 1485.93, 0.], np.float32)
KAT_SVD_TK = np.array([0.00111412, 0.00730956, -0.00647493, 0.999952, -10.4598, 4.74009, -0.762817, 1.00578], np.float32)


def test_power_method_kat(ctx, po, alg):             # testsICP.cpp:988-1087
    pm = alg.ICPPowerMethod(ctx); pm.init(); pm.write("D_IN_S", KAT_S); pm.write("D_IN_MEAN", KAT_MEANS); pm.run()
    got = pm.read()
    want, _ = po.power_method(KAT_S, KAT_MEANS)
    assert_bits_equal(got, want, "Tk")
    assert np.abs(got - KAT_SVD_TK).max() < 42000 * np.finfo(np.float32).eps     # the reference's SVD golden band


def test_power_method_random(ctx, po, alg):
    rng = np.random.default_rng(10)
    pm = alg.ICPPowerMethod(ctx); pm.init()
    for _ in range(20):
        R = rng.normal(size=(3, 3)); S3 = (R @ R.T) * 1e-3 + rng.normal(size=(3, 3)) * 1e-4
        if _ % 4 == 3:
            S3 = -S3                                  # exercises the negative-eigenvalue shift + restart
        S = np.r_[S3.reshape(-1), 0.0052, 0.0051].astype(np.float32)
        mu = np.r_[rng.uniform(-50, 50, 2), 1500, 0, rng.uniform(-50, 50, 2), 1490, 0].astype(np.float32)
        pm.write("D_IN_S", S); pm.write("D_IN_MEAN", mu); pm.run()
        want, _ = po.power_method(S, mu)
        assert_bits_equal(pm.read(), want, "Tk(random)")


def test_svd_solve(ctx, po, alg):                    # golden: testsICP.cpp:1042-1046
    sv = alg.ICPSVD(ctx); sv.init(); sv.write("D_IN_S", KAT_S); sv.write("D_IN_MEAN", KAT_MEANS); sv.run()
    Tk, Rk = sv.read()
    Tko, Rko = po.svd_solve(KAT_S, KAT_MEANS)
    assert_bits_equal(Tk, Tko, "svd Tk"); assert_bits_equal(Rk, Rko, "svd Rk")
    assert np.abs(Tk - KAT_SVD_TK).max() < 42000 * np.finfo(np.float32).eps
    rng = np.random.default_rng(12)
    for i in range(20):
        S3 = rng.normal(size=(3, 3)) * 1e-3
        if i % 5 == 4:
            S3[:, 2] = 0                               # rank deficient
        S = np.r_[S3.reshape(-1), 0.0052, 0.0051].astype(np.float32)
        sv.write("D_IN_S", S); sv.run()
        Tk, Rk = sv.read(); Tko, Rko = po.svd_solve(S, KAT_MEANS)
        assert_bits_equal(Tk, Tko, "svd Tk(random)"); assert_bits_equal(Rk, Rko, "svd Rk(random)")
        if i % 5 != 4:
            assert abs(np.linalg.det(Rk.astype(np.float64)) - 1) < 1e-4


@pytest.mark.parametrize("cols,rows", [(1024, 1024), (512, 3), (4096, 11), (76800, 11), (4, 2)])
def test_reduce_sum(ctx, po, alg, cols, rows):       # testsReduce.cpp:226
    rng = np.random.default_rng(13)
    a = rng.uniform(0, 1, (rows, cols)).astype(np.float32)
    r = alg.Reduce(ctx, "SUM"); r.init(cols, rows); r.write("D_IN", a); r.run()
    got = r.read()
    assert_bits_equal(got, po.reduce_sum_f(a), "reduce_sum")
    assert np.abs(got - a.astype(np.float64).sum(1)).max() < 42000 * np.finfo(np.float32).eps * max(1, cols / 1024)


def test_reduce_min_max(ctx, po, alg):               # testsReduce.cpp:64,145
    rng = np.random.default_rng(14)
    a = rng.uniform(0, 1, (1024, 1024)).astype(np.float32)
    r = alg.Reduce(ctx, "MIN"); r.init(1024, 1024); r.write("D_IN", a); r.run()
    assert_bits_equal(r.read(), po.reduce_min_f(a), "reduce_min")
    u = rng.integers(0, 2 ** 32, (1024, 1024), dtype=np.uint32)
    r = alg.Reduce(ctx, "MAX"); r.init(1024, 1024); r.write("D_IN", u); r.run()
    assert np.array_equal(r.read(), po.reduce_max_ui(u))
    r = alg.Reduce(ctx, "MIN"); r.init(6, 1)
    with pytest.raises(alg.ICPConfigError):
        r.run()


@pytest.mark.parametrize("cols,rows", [(1024, 1024), (1, 3), (1000, 7), (300, 1)])
def test_scan(ctx, po, alg, cols, rows):             # testsScan.cpp:65,150
    rng = np.random.default_rng(15)
    a = rng.integers(0, 10000, (rows, cols)).astype(np.int32)
    for cfg in ("INCLUSIVE", "EXCLUSIVE"):
        s = alg.Scan(ctx, cfg); s.init(cols, rows); s.write("D_IN", a); s.run()
        assert np.array_equal(s.read(), po.scan_i(a, cfg == "INCLUSIVE")), cfg


@pytest.mark.parametrize("n,nr,a", [(M, NR, 2e2), (M, 512, 2e2), (4096, 64, 1e-3), (1000, 12, 1.0), (65536, 1024, 2e2)])
def test_rbc_construct_and_search(ctx, po, alg, n, nr, a):
    """RBC build + 2-stage search on random 8-D clouds (collisions/ties included: duplicated points)."""
    rng = np.random.default_rng(16)
    X = rng_points(rng, n)
    X[n // 3] = X[n // 7]                              # exact duplicates => distance ties
    X[n // 2] = X[n // 7]
    R = X[rng.choice(n, nr, replace=False)].copy()
    R[nr // 2] = R[nr // 3]                            # duplicated representative => one empty list
    Q = rng_points(rng, n)
    Q[5] = X[n // 7]
    c = alg.RBCConstruct(ctx); c.init(n, nr, a); c.write("D_IN_X", X); c.write("D_IN_R", R); c.run()
    got = c.read(); want = po.rbc_construct(X, R, a)
    for k in ("rep_id", "N", "O", "perm"):
        assert np.array_equal(got[k], want[k]), k
    assert_bits_equal(got["Xp"], want["Xp"], "Xp")
    s = alg.RBCSearch(ctx); s.init(n, nr, a)
    s.write("D_IN_Q", Q); s.write("D_IN_R", R); s.write("D_IN_X_P", got["Xp"]); s.write("D_IN_O", got["O"]); s.write("D_IN_N", got["N"])
    s.run()
    g = s.read(); w = po.rbc_search(Q, R, a, want["Xp"], want["O"], want["N"])
    nonempty = want["N"][w["q_rep"][w["qperm"]]] > 0
    for k in ("q_rep", "qperm", "Nq", "Oq"):
        assert np.array_equal(g[k], w[k]), k
    assert np.array_equal(g["nn_id"][nonempty], w["nn_id"][nonempty]), "nn_id"
    assert_bits_equal(g["nn_dist"], w["nn_dist"], "nn_dist")
    assert_bits_equal(g["Qp"], w["Qp"], "Qp")
    assert_bits_equal(g["NN"][nonempty], w["NN"][nonempty], "NN")
    # the search is exact inside the chosen list
    d = ((g["Qp"][:64, None, :] - got["Xp"][None, :, :]) ** 2)
    assert (g["nn_id"][:64] < n).all() and d.shape[1] == n
