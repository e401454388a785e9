"""Parity of the BENCHMARKED configuration (what bench.py times): the batch engine with the sorted kernels B'/C', the exact
temporal pruning of stages 1 and 2 (settle), kernel D in the tail of C', concurrent slices, 40 fixed iterations, through
icp_batch_register, icp_batch_register_host and icp_batch_register_host_async / icp_batch_collect.

Every pair's pose is compared bit for bit with the oracle loop (restating src/ICP/algorithms.cpp:4670-4698 under the
fixed-count driver of algorithms.hpp:2487-2489) after 10, 20, 30 and 40 iterations -- the temporal pruning does most of
its skipping after iteration ~10 -- and the NN indices / sorted order of several pairs at the same checkpoints."""
import os

import numpy as np
import pytest

from util import assert_bits_equal

pytestmark = pytest.mark.gpu

M, NR, ITERS = 16384, 256, 40
CHECKPOINTS = (10, 20, 30, 40)


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


@pytest.fixture(scope="module")
def bench_batch(ctx, po, alg):
    """64 synthetic pairs generated exactly like bench.py's (icp_batch_synthesize, seed 5000), 4 concurrent slices, and
    the oracle's 40-iteration history of every pair."""
    from icp_b200 import synth
    n_pairs = 64
    for k in ("ICP_B200_SETTLE", "ICP_B200_FUSED", "ICP_B200_CMODE", "ICP_B200_AMODE", "ICP_B200_NNWALK"):
        assert k not in os.environ, f"{k} is set: this test must run the default (benchmarked) configuration"
    b = alg.ICPBatch(ctx, n_pairs, M, NR, a=2e2, c=1e-6)
    b.set_slices(4)
    base = ctx.upload(synth.base_landmarks())
    b.synthesize(base, 5000)
    ctx.sync()
    assert b.config()["QB"] == 1024 and b.cmode() in (2, 3) and b.slices() == 4, "not the benchmarked kernel configuration"
    hF = np.stack([b.debug("F", np.float32, (M, 8), pair=p) for p in range(n_pairs)])
    hM = np.stack([b.debug("M", np.float32, (M, 8), pair=p) for p in range(n_pairs)])
    detail = (0, 15, 16, 63)                     # first / last pair of a slice, slice boundary, last pair of the batch
    refs = []
    for p in range(n_pairs):
        r = po.icp_register(hF[p], hM[p], 128, 128, NR, a=2e2, c=1e-6, rot="power", weighted=True, fixed_iters=ITERS, dumps=True)
        keep = dict(T_hist=r["T_hist"][[k - 1 for k in CHECKPOINTS]].copy(), T=r["T"].copy())
        if p in detail:
            keep["nn_id"] = {k: r["nn_id_hist"][k - 1].copy() for k in CHECKPOINTS}
            keep["qperm"] = {k: r["qperm_hist"][k - 1].copy() for k in CHECKPOINTS}
        refs.append(keep)
    yield b, hF, hM, refs, detail
    b.close()


def test_register_matches_oracle_at_10_20_30_40(bench_batch, alg):
    b, hF, hM, refs, detail = bench_batch
    for ci, k in enumerate(CHECKPOINTS):
        b.register(k)                            # a registration always restarts from buildRBC
        T8 = b.read_poses()
        for p in range(b.n_pairs):
            assert_bits_equal(T8[p], refs[p]["T_hist"][ci], f"pose of pair {p} after {k} iterations")
        for p in detail:
            ids = b.debug("NN_ID", alg.DIST_ID, M, pair=p)["id"]
            assert np.array_equal(ids, refs[p]["nn_id"][k]), f"NN ids of pair {p} after {k} iterations"
            assert np.array_equal(b.debug("qperm", np.uint32, M, pair=p), refs[p]["qperm"][k]), f"qperm of pair {p} after {k} iterations"


def test_host_entries_match_oracle_at_40(ctx, bench_batch, alg):
    """The two host-buffer entries bench.py's e2e leg uses, on fresh device buffers (everything arrives through the entry)."""
    b, hF, hM, refs, _ = bench_batch
    want = np.stack([r["T"] for r in refs])
    b1 = alg.ICPBatch(ctx, b.n_pairs, M, NR, a=2e2, c=1e-6)
    b2 = alg.ICPBatch(ctx, b.n_pairs, M, NR, a=2e2, c=1e-6)
    b1.set_slices(4); b2.set_slices(4)
    assert_bits_equal(b1.register_host(hF, hM, ITERS, 0), want, "icp_batch_register_host, 40 iterations")
    # two alternating batches, the second one with the pairs in reverse order (different data per step)
    hF2, hM2 = hF[::-1].copy(), hM[::-1].copy()
    b1.register_host_async(hF, hM, ITERS, 0)
    b2.register_host_async(hF2, hM2, ITERS, 0)
    assert_bits_equal(b1.collect(), want, "icp_batch_register_host_async / collect, batch 1")
    assert_bits_equal(b2.collect(), want[::-1], "icp_batch_register_host_async / collect, batch 2 (reversed pairs)")
    b1.close(); b2.close()


@pytest.mark.parametrize("kind", ["incoherent", "ties"])
def test_adversarial_cloud_40_iterations(ctx, po, alg, kind):
    """One cloud built to break the pruning bounds through the benchmarked kernels for the full 40 iterations (SVD solve:
    defined for every input), every checkpoint bit-exact."""
    from test_assign_pruning_gpu import clouds
    n_pairs = 12
    data = [clouds(kind, seed=31 + p) for p in range(2)]
    refs = [po.icp_register(F, Mv, 128, 128, NR, a=2e2, c=1e-6, rot="svd", weighted=True, fixed_iters=ITERS, dumps=True) for F, Mv in data]
    b = alg.ICPBatch(ctx, n_pairs, M, NR, rot=0)
    assert b.config()["QB"] == 1024 and b.cmode() in (2, 3)
    b.upload(0, np.stack([data[p % 2][0] for p in range(n_pairs)]), np.stack([data[p % 2][1] for p in range(n_pairs)]))
    for k in CHECKPOINTS:
        b.register(k)
        T8 = b.read_poses()
        for p in range(n_pairs):
            ref = refs[p % 2]
            got, want = T8[p], ref["T_hist"][k - 1]
            assert np.array_equal(np.isnan(got), np.isnan(want)), f"NaN pattern of pose {p} after {k}"
            ok = ~np.isnan(want)
            assert_bits_equal(got[ok], want[ok], f"{kind}: pose {p} after {k} iterations")
            if p < 2:
                assert np.array_equal(b.debug("NN_ID", alg.DIST_ID, M, pair=p)["id"], ref["nn_id_hist"][k - 1]), f"{kind}: nn_id pair {p} after {k}"
    b.close()
