"""world_size-2 gloo test (CPU) of the N>1 host logic: block partition of independent frame pairs, pose gather in
pair order, max-over-ranks timing.  The per-pair compute is replaced by a deterministic stand-in: the product has no
CPU path, and none is needed to test the plumbing."""
import os
import socket

import numpy as np
import pytest

from icp_b200 import parallel


def fake_pose(i):
    return np.array([i, i + 0.25, -i, 1.0, 10 * i, 0.5 * i, 3.0, 1.0], np.float32)


@pytest.mark.parametrize("n,world", [(4096, 8), (10, 3), (7, 2), (1, 1), (5, 8)])
def test_pair_range_is_a_partition(n, world):
    seen = []
    for r in range(world):
        lo, hi = parallel.pair_range(n, world, r)
        assert 0 <= lo <= hi <= n
        seen += list(range(lo, hi))
        for i in range(lo, hi):
            assert parallel.pair_owner(i, n, world) == r
    assert seen == list(range(n))
    sizes = [parallel.pair_range(n, world, r)[1] - parallel.pair_range(n, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_pairs, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = parallel.pair_range(n_pairs, world, rank)
    local = np.stack([fake_pose(i) for i in range(lo, hi)]) if hi > lo else np.zeros((0, 8), np.float32)
    out = parallel.gather_poses(local, n_pairs, dist=dist)
    tmax = parallel.max_over_ranks(1.0 + rank, dist=dist)
    dist.barrier()
    if rank == 0:
        q.put((out, tmax))
    else:
        assert out is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [9, 16])
def test_gather_poses_world2_gloo(n_pairs):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out.shape == (n_pairs, 8)
    want = np.stack([fake_pose(i) for i in range(n_pairs)])
    assert np.array_equal(out, want)
    assert tmax == 2.0
