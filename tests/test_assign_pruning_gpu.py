"""Kernel A with triangle-inequality pruning (k_assign_tri) must return exactly what the exhaustive scan returns.
Adversarial inputs for the bound: clouds without any spatial coherence, massive ties / duplicated representatives,
degenerate sets, non-finite and overflowing coordinates, extreme metric weights -- each checked bit-exact against the
oracle (exhaustive strict-'<' scans), with the pruned (default) and the exhaustive (ICP_B200_AMODE=0) flavour."""
import os

import numpy as np
import pytest

from util import assert_bits_equal

pytestmark = pytest.mark.gpu

M, NR = 16384, 256


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


@pytest.fixture(params=["pruned", "pruned+nnwalk", "exhaustive"])
def amode(request):
    """Kernel A flavours: triangle-pruned stage 1 (default), + pruned stage-2 walk (opt-in), exhaustive stage 1."""
    old = {k: os.environ.get(k) for k in ("ICP_B200_AMODE", "ICP_B200_NNWALK")}
    os.environ.pop("ICP_B200_AMODE", None)
    os.environ.pop("ICP_B200_NNWALK", None)
    if request.param == "exhaustive":
        os.environ["ICP_B200_AMODE"] = "0"
    elif request.param == "pruned+nnwalk":
        os.environ["ICP_B200_NNWALK"] = "1"
    yield request.param
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


def pc8d(xyz, rgb):
    p = np.ones((len(xyz), 8), np.float32)
    p[:, :3] = xyz
    p[:, 4:7] = rgb
    return p


def clouds(kind, seed=5):
    rng = np.random.default_rng(seed)
    if kind == "incoherent":            # uniform in a box, random order: neighbouring indices are unrelated
        F = pc8d(rng.uniform(-2000, 2000, (M, 3)), rng.uniform(0, 1, (M, 3)))
        Mv = pc8d(rng.uniform(-2000, 2000, (M, 3)), rng.uniform(0, 1, (M, 3)))
    elif kind == "ties":                # coarse lattice: many exactly equal distances, duplicated points and representatives
        F = pc8d(rng.integers(0, 6, (M, 3)) * 100.0, rng.integers(0, 2, (M, 3)) * 0.5)
        Mv = pc8d(rng.integers(0, 6, (M, 3)) * 100.0 + 50.0, rng.integers(0, 2, (M, 3)) * 0.5)
    elif kind == "identical":           # every point the same: all distances 0, everything ties
        F = pc8d(np.full((M, 3), 123.0), np.full((M, 3), 0.25))
        Mv = F.copy()
    elif kind == "two_clusters":        # half of the representatives coincide
        F = pc8d(np.where(rng.uniform(size=(M, 1)) < 0.5, 0.0, 1000.0) + rng.normal(0, 1e-3, (M, 3)), rng.uniform(0, 1, (M, 3)))
        F[::2, :3] = 0.0
        Mv = pc8d(rng.uniform(-10, 1010, (M, 3)), rng.uniform(0, 1, (M, 3)))
    elif kind == "tiny":                # squares underflow: the absolute slack of the bound is exercised
        F = pc8d(rng.uniform(-1e-20, 1e-20, (M, 3)), rng.uniform(0, 1e-20, (M, 3)))
        Mv = pc8d(rng.uniform(-1e-20, 1e-20, (M, 3)), rng.uniform(0, 1e-20, (M, 3)))
    elif kind == "huge":                # squares overflow to +inf for part of the pairs
        F = pc8d(rng.uniform(-1e19, 1e19, (M, 3)) * (rng.uniform(size=(M, 1)) < 0.3) + rng.uniform(-100, 100, (M, 3)), rng.uniform(0, 1, (M, 3)))
        Mv = pc8d(rng.uniform(-1e19, 1e19, (M, 3)) * (rng.uniform(size=(M, 1)) < 0.3) + rng.uniform(-100, 100, (M, 3)), rng.uniform(0, 1, (M, 3)))
    elif kind == "nonfinite_points":    # NaN / inf in some moving and fixed points (not in the representatives' cells only)
        F = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        Mv = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        bad = rng.choice(M, 300, replace=False)
        Mv[bad[:100], 0] = np.nan
        Mv[bad[100:200], 1] = np.inf
        Mv[bad[200:], 5] = -np.inf
    elif kind == "nonfinite_reps":      # NaN / inf inside the representative set itself: the neighbour table is rejected
        F = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        Mv = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        F[(3 * 128 + 3)] = F[(3 * 128 + 3)] * np.float32(np.nan)          # representative 0
        F[(11 * 128 + 11), 2] = np.inf                                     # representative 17
    elif kind == "w_lanes_vary":        # homogeneous lanes not constant: the 8-lane distance path
        F = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        Mv = pc8d(rng.uniform(-500, 500, (M, 3)), rng.uniform(0, 1, (M, 3)))
        F[:, 3] = rng.uniform(0.5, 1.5, M); F[:, 7] = rng.uniform(0.5, 1.5, M)
        Mv[:, 3] = rng.uniform(0.5, 1.5, M); Mv[:, 7] = rng.uniform(0.5, 1.5, M)
    else:
        raise ValueError(kind)
    return F.astype(np.float32), Mv.astype(np.float32)


KINDS = ["incoherent", "ties", "identical", "two_clusters", "tiny", "huge", "nonfinite_points", "nonfinite_reps", "w_lanes_vary"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("a", [2e2, 1e-3, 1e6])
def test_assignments_and_nn_match_exhaustive_oracle(ctx, po, alg, amode, kind, a):
    if a != 2e2 and kind not in ("incoherent", "ties", "two_clusters"):
        pytest.skip("extreme alpha only on the clouds where it changes the neighbourhoods")
    F, Mv = clouds(kind)
    s = alg.ICPStep(ctx, 1, 1)
    s.init(M, NR, a, 1e-6)
    s.set_mode(alg.capi.MODE_FUSED)
    s.write(alg.capi.MEM_D_IN_F, F)
    s.write(alg.capi.MEM_D_IN_M, Mv)
    reps = po.get_reps(F, 128, 128, NR)
    want = po.rbc_construct(F, reps, a)
    T0 = np.array([0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    sr = po.rbc_search(po.transform_q(Mv, T0), reps, a, want["Xp"], want["O"], want["N"])
    for rep in range(2):                # second pass: seeds now come from the previous results
        s.reset()
        s.buildRBC()
        assert np.array_equal(s.debug("rep_id", np.uint32, M), want["rep_id"]), f"rep_id pass {rep}"
        assert np.array_equal(s.debug("N", np.uint32, NR), want["N"])
        assert np.array_equal(s.debug("perm", np.uint32, M), want["perm"])
        s.run(1)
        assert np.array_equal(s.debug("q_rep", np.uint32, M), sr["q_rep"]), f"q_rep pass {rep}"
        assert np.array_equal(s.debug("qperm", np.uint32, M), sr["qperm"])
        nnid = s.debug("NN_ID", alg.DIST_ID, M)
        assert np.array_equal(nnid["id"], sr["nn_id"]), f"nn_id pass {rep}"
        assert_bits_equal(nnid["dist"], sr["nn_dist"], "nn_dist")
    s.close()


@pytest.mark.parametrize("kind", ["incoherent", "ties", "two_clusters"])
def test_iterations_follow_the_oracle(ctx, po, alg, amode, kind):
    """Seeds are the previous iteration's representatives: several iterations, every pose bit-exact."""
    F, Mv = clouds(kind, seed=9)
    K = 4
    ref = po.icp_register(F, Mv, 128, 128, NR, a=2e2, c=1e-6, rot="svd", weighted=True, fixed_iters=K, dumps=True)
    s = alg.ICPStep(ctx, 0, 1)
    s.init(M, NR, 2e2, 1e-6)
    s.set_mode(alg.capi.MODE_FUSED)
    s.write(alg.capi.MEM_D_IN_F, F)
    s.write(alg.capi.MEM_D_IN_M, Mv)
    s.buildRBC()
    for k in range(K):
        s.run(1)
        assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, M)["id"], ref["nn_id_hist"][k]), f"nn_id it{k}"
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T_hist"][k], f"T it{k}")
    s.close()


def test_nn_walk_settles_queries_and_matches_a_full_registration(ctx, po, alg):
    """The opt-in stage-2 walk (anchored at last iteration's match) on the BASELINE workload: 12 iterations bit-exact,
    and it really does settle a large share of the queries (fewer executed than algorithmic stage-2 evaluations)."""
    from util import scene_pair
    F, Mv, _, _ = scene_pair(seed=21)
    K = 12
    ref = po.icp_register(F, Mv, 128, 128, NR, a=2e2, c=1e-6, rot="power", weighted=True, fixed_iters=K, dumps=True)
    old = os.environ.get("ICP_B200_NNWALK")
    os.environ["ICP_B200_NNWALK"] = "1"
    try:
        s = alg.ICPStep(ctx, 1, 1)
        s.init(M, NR, 2e2, 1e-6)
        s.set_mode(alg.capi.MODE_FUSED)
        s.write(alg.capi.MEM_D_IN_F, F)
        s.write(alg.capi.MEM_D_IN_M, Mv)
        s.set_count_evals(True)
        s.buildRBC()
        for k in range(K):
            s.run(1)
            nnid = s.debug("NN_ID", alg.DIST_ID, M)
            assert np.array_equal(nnid["id"], ref["nn_id_hist"][k]), f"nn_id it{k}"
            assert_bits_equal(s.debug("T", np.float32, 8), ref["T_hist"][k], f"T it{k}")
        e1, e2 = s.eval_counts()
        e1x, e2x = s.stage1_executed(), s.stage2_executed()
        assert e1 == K * M * NR and 0 < e1x < e1 // 4, (e1, e1x)
        assert 0 < e2x < e2, (e2, e2x)
        s.close()
    finally:
        if old is None:
            os.environ.pop("ICP_B200_NNWALK", None)
        else:
            os.environ["ICP_B200_NNWALK"] = old


def test_pruning_is_off_for_metric_weights_outside_the_proof(ctx, po, alg):
    """fg / fp outside [0, 1] (icp_step_set_metric): the pruned kernel must fall back to the exhaustive scan."""
    F, Mv = clouds("incoherent", seed=3)
    s = alg.ICPStep(ctx, 1, 1)
    s.init(M, NR, 2e2, 1e-6)
    if not hasattr(s, "set_metric"):
        pytest.skip("set_metric is not exposed by the Python mirror")
    s.set_mode(alg.capi.MODE_FUSED)
    s.set_metric(2.5, 0.75)
    s.write(alg.capi.MEM_D_IN_F, F)
    s.write(alg.capi.MEM_D_IN_M, Mv)
    s.buildRBC()
    reps = po.get_reps(F, 128, 128, NR)
    d = ((F[:, None, :4].astype(np.float32) - reps[None, :, :4]) ** 2)
    # exhaustive argmin in float32 with the oracle's op order
    g = ((d[..., 0] + d[..., 1]) + d[..., 2]) + d[..., 3]
    c = ((F[:, None, 4:].astype(np.float32) - reps[None, :, 4:]) ** 2)
    p = ((c[..., 0] + c[..., 1]) + c[..., 2]) + c[..., 3]
    dist = np.float32(2.5) * g + np.float32(0.75) * p
    assert np.array_equal(s.debug("rep_id", np.uint32, M), dist.argmin(1).astype(np.uint32))
    s.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("settle", ["1", "0"])
def test_batch_engine_on_adversarial_clouds(ctx, po, alg, kind, settle):
    """Throughput configuration (sorting kernel B', sorted kernel C' with the exact temporal pruning of stage 2, kernel D
    in its tail) on the clouds built to break bounds: every iteration's result must be the exhaustive oracle's.
    With settle=1 the runner-up bounds meet duplicates (gap 0), NaN / inf distances, underflowing and overflowing
    squares and poses that jump; settle=0 is the same kernels without the pruning."""
    old = os.environ.get("ICP_B200_SETTLE")
    os.environ["ICP_B200_SETTLE"] = settle
    try:
        n_pairs, K = 10, 5
        data = [clouds(kind, seed=5 + p) for p in range(3)]
        b = alg.ICPBatch(ctx, n_pairs, M, NR, rot=0)            # SVD solve: defined for every input
        assert b.config()["QB"] == 1024 and b.cmode() in (2, 3), "not the batch-mode configuration"
        b.upload(0, np.stack([data[p % 3][0] for p in range(n_pairs)]), np.stack([data[p % 3][1] for p in range(n_pairs)]))
        refs = [po.icp_register(F, Mv, 128, 128, NR, a=2e2, c=1e-6, rot="svd", weighted=True, fixed_iters=K, dumps=True) for F, Mv in data]
        for k in (1, 2, K):                                      # a registration always restarts from the build
            b.register(k)
            T8 = b.read_poses()
            for p in range(n_pairs):
                ref = refs[p % 3]
                assert np.array_equal(b.debug("NN_ID", alg.DIST_ID, M, pair=p)["id"], ref["nn_id_hist"][k - 1]), f"nn_id pair {p} after {k}"
                assert np.array_equal(b.debug("qperm", np.uint32, M, pair=p), ref["qperm_hist"][k - 1]), f"qperm pair {p} after {k}"
                got, want = T8[p], ref["T_hist"][k - 1]
                assert np.array_equal(np.isnan(got), np.isnan(want)), f"NaN pattern of pose {p} after {k}"
                ok = ~np.isnan(want)
                assert_bits_equal(got[ok], want[ok], f"pose {p} after {k} iterations")
        b.close()
    finally:
        if old is None:
            os.environ.pop("ICP_B200_SETTLE", None)
        else:
            os.environ["ICP_B200_SETTLE"] = old


def test_settle_skips_scans_on_the_baseline_workload(ctx, po, alg):
    """The temporal pruning really does settle queries on converging pairs (fewer executed than algorithmic stage-2
    evaluations) while every pose stays the oracle's."""
    from icp_b200 import synth
    old = os.environ.get("ICP_B200_BATCH_EVALS")
    os.environ["ICP_B200_BATCH_EVALS"] = "1"
    try:
        n_pairs, K = 10, 12
        b = alg.ICPBatch(ctx, n_pairs, M, NR)
        base = ctx.upload(synth.base_landmarks())
        b.synthesize(base, 9300)
        b.register(K)
        T8 = b.read_poses()
        for p in (0, 9):
            F = b.debug("F", np.float32, (M, 8), pair=p)
            Mv = b.debug("M", np.float32, (M, 8), pair=p)
            ref = po.icp_register(F, Mv, 128, 128, NR, fixed_iters=K, dumps=True)
            assert_bits_equal(T8[p], ref["T"], f"pose {p}")
            ev = b.debug("evals", np.uint64, 4, pair=p)
            assert int(ev[1]) == int(np.sum(ref["e2_hist"])), (ev, np.sum(ref["e2_hist"]))     # algorithmic count = the oracle's
            assert 0 < int(ev[3]) < int(ev[1]), ev                                             # executed < algorithmic
        b.close()
    finally:
        if old is None:
            os.environ.pop("ICP_B200_BATCH_EVALS", None)
        else:
            os.environ["ICP_B200_BATCH_EVALS"] = old
