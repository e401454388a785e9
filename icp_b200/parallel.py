"""Multi-GPU host logic: a single registration does not shard (3-4 global reductions per iteration on a
16 K-point problem), so scaling is by partitioning INDEPENDENT frame pairs over the ranks
(one process per GPU).  No collective sits on the hot path; only the final poses (8 floats per pair) are gathered.

torch.distributed is plumbing here (NCCL on the GPU box, gloo in the CPU tests); nothing in this file touches
the point data.
"""
import numpy as np


def pair_range(n_pairs_total, world_size, rank):
    """Contiguous block partition: rank r owns pairs [lo, hi).  Remainders go to the lowest ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n_pairs_total, world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def pair_owner(pair_index, n_pairs_total, world_size):
    """Inverse of pair_range: which rank owns a pair."""
    base, rem = divmod(n_pairs_total, world_size)
    cut = rem * (base + 1)
    if pair_index < cut:
        return pair_index // (base + 1)
    return rem + (pair_index - cut) // max(base, 1)


def gather_poses(local_poses, n_pairs_total, dist=None, device=None, dst=0):
    """Gather the per-rank pose blocks (k_r x 8 float32, in pair order) on rank `dst`.

    Returns the (n_pairs_total x 8) array on dst, None elsewhere.  With dist=None (single process) it is the identity.
    Uneven blocks are padded to the largest block for the collective and trimmed afterwards."""
    local_poses = np.ascontiguousarray(local_poses, np.float32).reshape(-1, 8)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        assert len(local_poses) == n_pairs_total
        return local_poses
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = pair_range(n_pairs_total, world, rank)
    assert len(local_poses) == hi - lo, (len(local_poses), lo, hi)
    kmax = -(-n_pairs_total // world)
    pad = np.zeros((kmax, 8), np.float32)
    pad[: hi - lo] = local_poses
    t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank != dst:
        return None
    blocks = []
    for r in range(world):
        a, b = pair_range(n_pairs_total, world, r)
        blocks.append(out[r].cpu().numpy()[: b - a])
    return np.concatenate(blocks, 0)


def max_over_ranks(value, dist=None, device=None):
    """Timing rule: a multi-GPU number is the MAX over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
