"""GPU parity tests of the registration engine (ICPStep / ICP / batch) through the C ABI against the oracle:
representative assignments, list permutations, NN indices bit-exact at every iteration; every reduction,
incremental and accumulated pose bit-exact (stricter than north_star's 1e-5), in both execution modes."""
import numpy as np
import pytest

from util import assert_bits_equal, scene_pair

pytestmark = pytest.mark.gpu

M, NR = 16384, 256
ROT = {"power": 1, "svd": 0}


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


@pytest.fixture(scope="module")
def pair():
    return scene_pair(seed=11)


def make_step(alg, ctx, F, M_, rot, weighted, mode, m=M, nr=NR, a=2e2, c=1e-6, lm=(0, 0), cls=None):
    s = (cls or alg.ICPStep)(ctx, ROT[rot], 1 if weighted else 0)
    if cls is alg.ICP:
        s.init(m, nr, a, c, 40, 0.001, 0.01, lm[0], lm[1])
    else:
        s.init(m, nr, a, c, lm[0], lm[1])
    s.set_mode(mode)
    s.write(alg.capi.MEM_D_IN_F, F)
    s.write(alg.capi.MEM_D_IN_M, M_)
    return s


@pytest.mark.parametrize("mode", [0, 1])
def test_build_rbc(ctx, po, alg, pair, mode):
    F, M_, _, _ = pair
    s = make_step(alg, ctx, F, M_, "power", True, mode)
    s.buildRBC()
    reps = po.get_reps(F, 128, 128, NR)
    want = po.rbc_construct(F, reps, 2e2)
    assert_bits_equal(s.debug("reps", np.float32, (NR, 8)), reps, "reps")
    assert np.array_equal(s.debug("rep_id", np.uint32, M), want["rep_id"])
    assert np.array_equal(s.debug("N", np.uint32, NR), want["N"])
    assert np.array_equal(s.debug("O", np.uint32, NR), want["O"])
    assert np.array_equal(s.debug("perm", np.uint32, M), want["perm"])
    assert_bits_equal(s.debug("Xp", np.float32, (M, 8)), want["Xp"], "Xp")
    assert want["N"].sum() == M and want["N"].min() > 0
    s.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("rot,weighted", [("power", True), ("svd", True), ("power", False), ("svd", False)])
def test_step_by_step(ctx, po, alg, pair, mode, rot, weighted):
    """Every intermediate of the first iterations, one ICPStep::run at a time (the ICPSBS use case)."""
    F, M_, _, _ = pair
    K = 3
    s = make_step(alg, ctx, F, M_, rot, weighted, mode)
    s.buildRBC()
    ref = po.icp_register(F, M_, 128, 128, NR, a=2e2, c=1e-6, rot=rot, weighted=weighted, fixed_iters=K, dumps=True)
    for k in range(K):
        s.run(1)
        assert np.array_equal(s.debug("qperm", np.uint32, M), ref["qperm_hist"][k]), f"qperm it{k}"
        nnid = s.debug("NN_ID", alg.DIST_ID, M)
        assert np.array_equal(nnid["id"], ref["nn_id_hist"][k]), f"nn_id it{k}"
        if weighted:
            sw = s.debug("sum_w", np.float64, 1)[0]
            assert np.float64(sw).view(np.uint64) == np.float64(ref["sumw_hist"][k]).view(np.uint64), (k, sw, ref["sumw_hist"][k])
        assert_bits_equal(s.debug("mean", np.float32, 8), ref["mean_hist"][k], f"mean it{k}")
        assert_bits_equal(s.debug("S", np.float32, 11), ref["S_hist"][k], f"S it{k}")
        assert_bits_equal(s.debug("Tk", np.float32, 8), ref["Tk_hist"][k], f"Tk it{k}")
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T_hist"][k], f"T it{k}")
    st = s.state()
    assert st["k"] == K
    s.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_first_iteration_details(ctx, po, alg, pair, mode):
    """Stage-level outputs of iteration 1 against the oracle's stage functions."""
    F, M_, _, _ = pair
    s = make_step(alg, ctx, F, M_, "power", True, mode)
    s.buildRBC(); s.run(1)
    reps = po.get_reps(F, 128, 128, NR)
    rb = po.rbc_construct(F, reps, 2e2)
    T0 = np.array([0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    Mt = po.transform_q(M_, T0)
    sr = po.rbc_search(Mt, reps, 2e2, rb["Xp"], rb["O"], rb["N"])
    assert np.array_equal(s.debug("q_rep", np.uint32, M), sr["q_rep"])
    assert np.array_equal(s.debug("Nq", np.uint32, NR), sr["Nq"])
    assert np.array_equal(s.debug("Oq", np.uint32, NR), sr["Oq"])
    nnid = s.debug("NN_ID", alg.DIST_ID, M)
    assert np.array_equal(nnid["id"], sr["nn_id"])
    assert_bits_equal(nnid["dist"], sr["nn_dist"], "nn_dist")
    W, sw = po.weights(sr["nn_dist"])
    assert_bits_equal(s.debug("W", np.float32, M), W, "W")
    if mode == 0:
        assert_bits_equal(s.debug("Qp", np.float32, (M, 8)), sr["Qp"], "Qp")
        assert_bits_equal(s.debug("NN", np.float32, (M, 8)), sr["NN"], "NN")
        mean = po.mean_weighted(sr["NN"], sr["Qp"], W, sw)
        DF, DM = po.devs(sr["NN"], sr["Qp"], mean)
        assert_bits_equal(s.debug("DF", np.float32, (M, 4)), DF, "DF")
        assert_bits_equal(s.debug("DM", np.float32, (M, 4)), DM, "DM")
    else:
        assert_bits_equal(s.debug("fxyz", np.float32, (3, M)), sr["NN"][:, :3].T, "fxyz")
        assert_bits_equal(s.debug("mxyz", np.float32, (3, M)), sr["Qp"][:, :3].T, "mxyz")
    s.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("rot", ["power", "svd"])
def test_registration_40_iterations(ctx, po, alg, pair, mode, rot):
    """BASELINE config: |F|=|M|=16384, |R|=256, a=2e2, c=1e-6, 40 fixed iterations; 4x4 pose within 1e-5."""
    F, M_, R_gt, t_gt = pair
    s = make_step(alg, ctx, F, M_, rot, True, mode)
    s.buildRBC(); s.run(40)
    ref = po.icp_register(F, M_, 128, 128, NR, rot=rot, weighted=True, fixed_iters=40, dumps=True)
    T16 = s.pose_matrix()
    assert np.abs(T16 - ref["T16"]).max() <= 1e-5, np.abs(T16 - ref["T16"]).max()
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "T after 40")
    assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, M)["id"], ref["nn_id_hist"][39])
    st = s.state()
    assert st["k"] == 40
    # sanity: the estimate moves towards the ground truth
    err0 = np.linalg.norm(t_gt)
    assert np.linalg.norm(T16[:3, 3] - t_gt) < err0
    s.close()


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_launch_variants_agree(ctx, po, alg, pair, variant):
    F, M_, _, _ = pair
    ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=7)
    for mode in (0, 1):
        if variant == 3 and mode == 0:
            continue                    # the persistent cooperative kernel is a fused-mode engine
        s = make_step(alg, ctx, F, M_, "power", True, mode)
        s.buildRBC(); s.run(7, variant=variant)
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], f"T variant {variant} mode {mode}")
        assert s.state()["k"] == 7
        s.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_icp_run_converges_like_the_reference_loop(ctx, po, alg, mode):
    """ICP::run with the thresholded check() on the device: same iteration count, same pose."""
    F, M_, _, _ = scene_pair(seed=12, deg=0.4, t=(2.0, -1.0, 1.5))
    for thr in ((0.001, 0.01), (0.05, 0.5)):
        icp = make_step(alg, ctx, F, M_, "power", True, mode, cls=alg.ICP)
        icp.angle_threshold, icp.translation_threshold = thr
        icp.buildRBC()
        k = icp.run()
        ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=0, max_iterations=40, angle_thr=thr[0], trans_thr=thr[1])
        assert k == ref["k"], (thr, k, ref["k"])
        assert_bits_equal(icp.debug("T", np.float32, 8), ref["T"], "T at convergence")
        icp.close()


@pytest.mark.parametrize("knob", ["ICP_B200_FASTD=0", "ICP_B200_WIDED=1", "ICP_B200_SF=8", "ICP_B200_SF=9", "ICP_B200_SF=16", "ICP_B200_QG=112", "ICP_B200_QB=512", "ICP_B200_TD=256", "ICP_B200_CMODE=2", "ICP_B200_CMODE=3", "ICP_B200_CMODE=0"])
def test_execution_knobs_do_not_change_results(ctx, po, alg, pair, knob):
    """Every tuning knob only changes how the work is laid out (generic kernel-D path instead of the shared-memory one,
    lanes per point in the exhaustive pass, CTA sizes): the poses stay bit-identical to the oracle's."""
    import os
    F, M_, _, _ = pair
    ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=6)
    k, v = knob.split("=")
    old = os.environ.get(k)
    os.environ[k] = v
    try:
        for rot in ("power", "svd"):
            s = make_step(alg, ctx, F, M_, rot, True, 1)
            s.buildRBC(); s.run(6)
            if rot == "power":
                assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], f"T with {knob}")
            assert s.state()["k"] == 6
            s.close()
        b = alg.ICPBatch(ctx, 3, M, NR)
        b.upload(0, np.stack([F, F, F]), np.stack([M_, M_, M_]))
        b.register(6)
        T8 = b.read_poses()
        for p in range(3):
            assert_bits_equal(T8[p], ref["T"], f"batch pose {p} with {knob}")
        b.close()
    finally:
        if old is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = old


def test_rebuild_and_rerun_is_reproducible(ctx, alg, pair):
    F, M_, _, _ = pair
    s = make_step(alg, ctx, F, M_, "power", True, 1)
    out = []
    for _ in range(2):
        s.reset(); s.buildRBC(); s.run(10)
        out.append(s.debug("T", np.float32, 8))
    assert_bits_equal(out[0], out[1], "rerun")
    s.close()


@pytest.mark.parametrize("m,nr,lm", [(65536, 512, (256, 256)), (65536, 1024, (256, 256)), (307200, 512, (640, 480)), (307200, 1024, (640, 480))])
def test_scaled_configs(ctx, po, alg, m, nr, lm):
    """BASELINE config 4 (beyond what the reference classes can run): oracle-only parity, 2 iterations."""
    from icp_b200 import synth
    F = synth.grid_cloud(*lm)
    F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
    ref = po.icp_register(F2, M_, lm[0], lm[1], nr, fixed_iters=2, dumps=True)
    for mode in (0, 1):
        s = make_step(alg, ctx, F2, M_, "power", True, mode, m=m, nr=nr, lm=lm)
        s.buildRBC(); s.run(2)
        assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, m)["id"], ref["nn_id_hist"][1]), f"nn_id mode {mode}"
        assert_bits_equal(s.debug("S", np.float32, 11), ref["S_hist"][1], f"S mode {mode}")
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], f"T mode {mode}")
        s.close()


def test_config_errors(ctx, alg):
    s = alg.ICPStep(ctx, 1, 1)
    with pytest.raises(alg.ICPConfigError):
        s.init(0, 256)
    with pytest.raises(alg.ICPConfigError):
        s.init(16384, 0)
    with pytest.raises(alg.ICPConfigError):
        s.init(16384, 256, a=0.0)
    s.close()


def test_batch_matches_single_engine(ctx, po, alg):
    """Throughput mode: every pair of a batch gets exactly the pose the single-pair engine / oracle computes."""
    from icp_b200 import synth
    n_pairs, K = 6, 5
    b = alg.ICPBatch(ctx, n_pairs, M, NR)
    base = ctx.upload(synth.base_landmarks())
    b.synthesize(base, 5000)
    b.register(K)
    T8 = b.read_poses()
    for p in (0, 3, 5):
        F = b.debug("F", np.float32, (M, 8), pair=p)
        M_ = b.debug("M", np.float32, (M, 8), pair=p)
        ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=K)
        assert_bits_equal(T8[p], ref["T"], f"batch pose {p}")
    # different pairs must differ
    assert not np.array_equal(T8[0], T8[1])
    b.close()


@pytest.mark.parametrize("knob", [None, "ICP_B200_CMODE=2", "ICP_B200_CMODE=1", "ICP_B200_CMODE=0", "ICP_B200_QG=512", "ICP_B200_QI=8", "ICP_B200_AMODE=0",
                                  "ICP_B200_SPAN_PTS=192", "ICP_B200_SPAN_PTS=64", "ICP_B200_SETTLE=0", "ICP_B200_DRING=0", "ICP_B200_FUSED=0"])
def test_batch_mode_kernels_match_oracle(ctx, po, alg, knob):
    """Throughput configuration (>= 10 pairs on a 148-SM part select the batch-mode kernels: 1024-query chunks, the
    sorted B'/C' flavour, 512-thread kernel D): every pose equals the oracle's, whichever kernel-C flavour runs."""
    import os
    from icp_b200 import synth
    n_pairs, K = 12, 4
    if knob:
        k, v = knob.split("=")
        old = os.environ.get(k)
        os.environ[k] = v
    try:
        b = alg.ICPBatch(ctx, n_pairs, M, NR)
        base = ctx.upload(synth.base_landmarks())
        b.synthesize(base, 8200)
        b.register(K)
        T8 = b.read_poses()
        assert b.config()["QB"] == 1024, "not the batch-mode configuration"
        for p in range(n_pairs):
            F = b.debug("F", np.float32, (M, 8), pair=p)
            M_ = b.debug("M", np.float32, (M, 8), pair=p)
            ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=K, dumps=True)
            assert_bits_equal(T8[p], ref["T"], f"batch-mode pose {p} ({knob})")
            if p in (0, 7):
                ids = b.debug("NN_ID", alg.DIST_ID, M, pair=p)["id"]
                assert np.array_equal(ids, ref["nn_id_hist"][K - 1]), f"NN ids of pair {p} ({knob})"
                assert np.array_equal(b.debug("qperm", np.uint32, M, pair=p), ref["qperm_hist"][K - 1]), f"qperm of pair {p} ({knob})"
        b.close()
    finally:
        if knob:
            if old is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old


def test_bind_rejects_misaligned_point_buffers(ctx, alg):
    """The fused kernels move points as 256-bit requests: icp_step_bind refuses point buffers that are not 32-byte aligned
    (every cudaMalloc / icp_malloc pointer is) instead of faulting later."""
    from icp_b200 import capi
    buf = ctx.alloc(M * 32 + 64)
    s = alg.ICPStep(ctx, capi.ROT_POWER_METHOD, capi.W_WEIGHTED)
    with pytest.raises(Exception):
        s.bind(capi.MEM_D_IN_F, buf.ptr + 16)
    s.bind(capi.MEM_D_IN_F, buf.ptr + 32)
    s.close()
    buf.free()


def test_batch_register_host_sliced(ctx, po, alg):
    """Host-buffer entry (sliced upload overlapped with compute): identical poses to register() on resident data and to
    the oracle, for slice counts that do and do not divide the batch, repeated calls (cached slice graphs) included."""
    from icp_b200 import synth
    n_pairs, K = 7, 4
    b = alg.ICPBatch(ctx, n_pairs, M, NR)
    base = ctx.upload(synth.base_landmarks())
    b.synthesize(base, 7100)
    b.register(K)
    want = b.read_poses()
    hF = np.stack([b.debug("F", np.float32, (M, 8), pair=p) for p in range(n_pairs)])
    hM = np.stack([b.debug("M", np.float32, (M, 8), pair=p) for p in range(n_pairs)])
    for n_slices in (2, 7):                        # register() as concurrent slices on resident data
        b.set_slices(n_slices)
        assert b.slices() == n_slices
        b.register(K)
        assert_bits_equal(b.read_poses(), want, f"register, {n_slices} concurrent slices")
    b.close()
    b = alg.ICPBatch(ctx, n_pairs, M, NR)          # fresh device buffers: everything must arrive through the host entry
    for n_slices in (1, 3, 7, 3, 100, 0):
        got = b.register_host(hF, hM, K, n_slices)
        assert_bits_equal(got, want, f"register_host, {n_slices} slices")
    ref = po.icp_register(hF[6], hM[6], 128, 128, NR, fixed_iters=K)
    assert_bits_equal(got[6], ref["T"], "register_host pose 6 vs oracle")
    # asynchronous pair on two alternating batches (what a streaming caller / bench.py's e2e leg does)
    b2 = alg.ICPBatch(ctx, n_pairs, M, NR)
    hF2, hM2 = hF[::-1].copy(), hM[::-1].copy()                # a second, different step: the pairs in reverse order
    b.register_host_async(hF, hM, K, 3)
    b2.register_host_async(hF2, hM2, K, 2)
    with pytest.raises(ValueError):
        b.register_host_async(hF, hM, K)                        # pending: must collect first
    assert_bits_equal(b.collect(), want, "async batch 1")
    assert_bits_equal(b2.collect(), want[::-1], "async batch 2")
    with pytest.raises(ValueError):
        b.collect()
    assert_bits_equal(b.register_host(hF, hM, K, 2), want, "blocking entry after the asynchronous one")
    b2.close()
    b.close()


def test_batch_upload_path(ctx, po, alg, pair):
    F, M_, _, _ = pair
    b = alg.ICPBatch(ctx, 2, M, NR)
    b.upload(0, np.stack([F, F]), np.stack([M_, M_]))
    b.register(4)
    T8, T16 = b.read_poses(want_T16=True)
    ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=4)
    assert_bits_equal(T8[0], ref["T"], "upload pose 0"); assert_bits_equal(T8[1], ref["T"], "upload pose 1")
    assert np.abs(T16[0] - ref["T16"]).max() <= 1e-5
    b.close()


@pytest.mark.parametrize("rot,weighted", [("power", True), ("svd", True), ("power", False), ("svd", False)])
def test_persistent_engine_matches_oracle(ctx, po, alg, pair, rot, weighted, monkeypatch):
    """The persistent cooperative kernel (one launch per run call, software grid barriers between the phases) as the engine of
    ICPStep::run and ICP::run: 40 fixed iterations and the thresholded loop, bit-exact poses, NN ids of the last iteration."""
    monkeypatch.setenv("ICP_B200_ENGINE", "persistent")
    F, M_, _, _ = pair
    s = make_step(alg, ctx, F, M_, rot, weighted, 1)
    s.buildRBC(); s.run(3); s.run(37)                      # two launches: state carried in device memory
    ref = po.icp_register(F, M_, 128, 128, NR, rot=rot, weighted=weighted, fixed_iters=40, dumps=True)
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "T after 40 (persistent engine)")
    assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, M)["id"], ref["nn_id_hist"][39])
    assert s.state()["k"] == 40
    s.close()
    F2, M2, _, _ = scene_pair(seed=12, deg=0.4, t=(2.0, -1.0, 1.5))
    icp = make_step(alg, ctx, F2, M2, rot, weighted, 1, cls=alg.ICP)
    icp.angle_threshold, icp.translation_threshold = 0.05, 0.5
    icp.buildRBC()
    k = icp.run()
    ref = po.icp_register(F2, M2, 128, 128, NR, rot=rot, weighted=weighted, fixed_iters=0, max_iterations=40, angle_thr=0.05, trans_thr=0.5)
    assert k == ref["k"], (k, ref["k"])
    assert_bits_equal(icp.debug("T", np.float32, 8), ref["T"], "T at convergence (persistent engine)")
    icp.close()


def test_persistent_engine_other_sizes(ctx, po, alg, monkeypatch):
    """65536 landmarks / 512 representatives (512-point chunks, generic kernel-D path inside the cluster) and a small set."""
    from icp_b200 import synth
    monkeypatch.setenv("ICP_B200_ENGINE", "persistent")
    F = synth.grid_cloud(256, 256)
    F2, M_, _, _ = synth.known_transform_pair(seed=78, deg=2.0, t=(10, -5, 8), F=F)
    ref = po.icp_register(F2, M_, 256, 256, 512, fixed_iters=3, dumps=True)
    s = make_step(alg, ctx, F2, M_, "power", True, 1, m=65536, nr=512, lm=(256, 256))
    s.buildRBC(); s.run(3)
    assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, 65536)["id"], ref["nn_id_hist"][2])
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "T, 65536 / 512")
    s.close()


def test_single_engine_temporal_pruning(ctx, po, alg, pair, monkeypatch):
    """The exact temporal pruning of the grouped kernel C (default for single registrations of >= 32768 points, forced here at
    16384): split run calls (bounds are re-trusted only inside a call), a replaced moving set between calls, every pose and
    the NN ids bit-exact, and the pruning really skips scans."""
    monkeypatch.setenv("ICP_B200_SETTLE", "1")
    F, M_, _, _ = pair
    ref = po.icp_register(F, M_, 128, 128, NR, fixed_iters=20, dumps=True)
    s = make_step(alg, ctx, F, M_, "power", True, 1)
    s.set_count_evals(True)
    s.buildRBC(); s.run(12); s.run(1); s.run(7)
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "T after 12 + 1 + 7 iterations")
    assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, M)["id"], ref["nn_id_hist"][19])
    e1, e2 = s.eval_counts()
    assert e2 == int(np.sum(ref["e2_hist"])) and 0 < s.stage2_executed() < e2
    F2, M2, _, _ = scene_pair(seed=13, deg=1.0, t=(4.0, 2.0, -3.0))
    s.reset(); s.write(alg.capi.MEM_D_IN_M, M2); s.run(6)            # new moving set, same RBC: stale bounds must not be used
    ref2 = po.icp_register(F, M2, 128, 128, NR, fixed_iters=6)
    assert_bits_equal(s.debug("T", np.float32, 8), ref2["T"], "T after replacing the moving set")
    s.close()


@pytest.mark.parametrize("nr", [64, 512, 1024])
def test_batch_engine_other_representative_counts(ctx, po, alg, nr):
    """Throughput configuration with |R| != 256 (the lane-order permutation, the shared-memory sort and the sorted kernel C'
    size their shared memory by |R|): poses and NN ids against the oracle."""
    from icp_b200 import synth
    n_pairs, K = 10, 4
    b = alg.ICPBatch(ctx, n_pairs, M, nr)
    base = ctx.upload(synth.base_landmarks())
    b.synthesize(base, 8300 + nr)
    b.register(K)
    T8 = b.read_poses()
    for p in (0, 4, 9):
        F = b.debug("F", np.float32, (M, 8), pair=p)
        M_ = b.debug("M", np.float32, (M, 8), pair=p)
        ref = po.icp_register(F, M_, 128, 128, nr, fixed_iters=K, dumps=True)
        assert np.array_equal(b.debug("NN_ID", alg.DIST_ID, M, pair=p)["id"], ref["nn_id_hist"][K - 1]), f"NN ids of pair {p}, |R| = {nr}"
        assert_bits_equal(T8[p], ref["T"], f"pose {p}, |R| = {nr}")
    b.close()


def test_large_single_registration_is_reproducible(ctx, po, alg):
    """One 307200-point registration spans several waves of kernel-A CTAs: the lane-order permutation of a chunk must only be
    trusted by the NEXT launch, and by a whole CTA at once (regression test of two races found in round 2).  Repeated
    registrations on one engine give the oracle's pose every time."""
    from icp_b200 import synth
    F = synth.grid_cloud(640, 480)
    F2, M_, _, _ = synth.known_transform_pair(seed=77, deg=2.0, t=(10, -5, 8), F=F)
    ref = po.icp_register(F2, M_, 640, 480, 512, fixed_iters=5)
    s = make_step(alg, ctx, F2, M_, "power", True, 1, m=307200, nr=512, lm=(640, 480))
    for rep in range(8):
        s.reset(); s.buildRBC(); s.run(5)
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], f"repetition {rep}")
    s.close()
