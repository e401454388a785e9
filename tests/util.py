"""Shared helpers of the test-suite."""
import numpy as np


def bits(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        return a.view(np.uint32)
    if a.dtype == np.float64:
        return a.view(np.uint64)
    return a


def assert_bits_equal(got, want, what=""):
    got = np.ascontiguousarray(got)
    want = np.ascontiguousarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    g, w = bits(got), bits(want)
    if not np.array_equal(g, w):
        bad = np.flatnonzero(g.reshape(-1) != w.reshape(-1))
        i = bad[0]
        raise AssertionError(f"{what}: {bad.size}/{g.size} elements differ; first at flat index {i}: "
                             f"got {got.reshape(-1)[i]!r} want {want.reshape(-1)[i]!r}")


def rng_points(rng, n, lo=0.0, hi=1.0):
    """n random pc8d-like points with lanes 3 and 7 = 1."""
    p = rng.uniform(lo, hi, (n, 8)).astype(np.float32)
    p[:, 3] = 1.0
    p[:, 7] = 1.0
    return p


def scene_pair(seed=11, m=16384, deg=3.0, t=(20.0, -15.0, 10.0)):
    """Seeded landmark pair with a known transform (config 3 style)."""
    from icp_b200 import synth
    F, M, R, tt = synth.known_transform_pair(seed=seed, deg=deg, t=t)
    assert len(F) == m
    return F, M, R, tt
