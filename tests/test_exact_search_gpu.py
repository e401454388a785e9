"""Exact random-ball-cover search (icp_rbc_search_exact) against the brute-force oracle: the nearest neighbour
over the WHOLE database, smallest distance and lowest list position among ties, bit-exact -- on scene-like clouds, clouds
built to break the triangle bound (ties, duplicated representatives / empty lists, underflow / overflow, NaN), a database
larger than the query set (frame-to-model) and extreme metric weights.  The pruning must also really prune."""
import numpy as np
import pytest

from util import assert_bits_equal, rng_points, scene_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


def run_exact(ctx, po, alg, X, R, Q, a):
    n, nr, m = len(X), len(R), len(Q)
    c = alg.RBCConstruct(ctx); c.init(n, nr, a); c.write("D_IN_X", X); c.write("D_IN_R", R); c.run()
    db = c.read()
    s = alg.RBCSearchExact(ctx); s.init(m, n, nr, a)
    s.write("D_IN_Q", Q); s.write("D_IN_R", R); s.write("D_IN_X_P", db["Xp"]); s.write("D_IN_O", db["O"]); s.write("D_IN_N", db["N"])
    s.run()
    got = s.read()
    want = po.nearest_exact(Q, db["Xp"], a)
    assert np.array_equal(got["nn_id"], want["nn_id"]), f"{np.count_nonzero(got['nn_id'] != want['nn_id'])} NN positions differ"
    assert_bits_equal(got["nn_dist"], want["nn_dist"], "nn_dist")
    assert_bits_equal(got["NN"], db["Xp"][want["nn_id"]], "NN points")
    return got, db


def test_scene_frame_to_frame_and_pruning_rate(ctx, po, alg):
    """The ICP workload itself: 16384 queries against the 16384-point fixed set, 256 representatives."""
    F, M_, _, _ = scene_pair(seed=61)
    R = po.get_reps(F, 128, 128, 256)
    got, db = run_exact(ctx, po, alg, F, R, M_, 2e2)
    brute = len(M_) * len(F)
    assert got["evals"] < brute // 8, (got["evals"], brute)        # the cover excludes most lists
    # the one-shot search of the ICP pipeline is a lower-quality answer: never closer than the exact one
    one = po.rbc_search(M_, R, 2e2, db["Xp"], db["O"], db["N"])
    exact_by_query = got["nn_dist"][one["qperm"]]
    assert (exact_by_query <= one["nn_dist"]).all() and (exact_by_query < one["nn_dist"]).any()


def test_frame_to_model(ctx, po, alg):
    """Database (model) four times larger than the frame that is matched against it."""
    from icp_b200 import synth
    model = synth.grid_cloud(256, 256)
    frame = scene_pair(seed=62)[1][::4].copy()
    R = po.get_reps(model, 256, 256, 512)
    got, _ = run_exact(ctx, po, alg, model, R, frame, 2e2)
    assert got["evals"] < len(frame) * len(model) // 8


@pytest.mark.parametrize("kind,a", [("ties", 2e2), ("two_clusters", 2e2), ("identical", 2e2), ("tiny", 2e2), ("huge", 2e2),
                                    ("nonfinite_points", 2e2), ("nonfinite_reps", 2e2), ("w_lanes_vary", 2e2),
                                    ("incoherent", 1e-3), ("incoherent", 1e6), ("ties", 1e6)])
def test_adversarial_clouds(ctx, po, alg, kind, a):
    from test_assign_pruning_gpu import clouds
    F, Mv = clouds(kind, seed=71)
    n = 4096
    X, Q = F[:n].copy(), Mv[:1024].copy()
    R = X[::16][:256].copy()
    if kind in ("ties", "two_clusters"):
        R[7] = R[3]                                     # duplicated representative: one empty list
    run_exact(ctx, po, alg, X, R, Q, a)


def test_random_cloud_with_duplicates_small_sizes(ctx, po, alg):
    rng = np.random.default_rng(5)
    for n, nr, m in ((1000, 12, 333), (64, 4, 64), (5000, 100, 7)):
        X = rng_points(rng, n)
        X[n // 3] = X[n // 7]
        R = X[rng.choice(n, nr, replace=False)].copy()
        Q = rng_points(rng, m)
        Q[0] = X[n // 7]
        run_exact(ctx, po, alg, X, R, Q, 1.0)


def test_config_errors(ctx, alg):
    s = alg.RBCSearchExact(ctx)
    s.init(16, 16, 4, 0.0)
    with pytest.raises(alg.ICPConfigError):
        s.run()
