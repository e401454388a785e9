"""Single-process multi-GPU entry (icp_multi_*: contiguous pair blocks, one host thread + context per device, no collective)
and two contexts on different devices inside one process (every entry point sets / restores the device it works on).
The 2-GPU cases skip on a single-GPU box; the 1-device cases run everywhere there is a GPU."""
import numpy as np
import pytest

from util import assert_bits_equal, scene_pair

pytestmark = pytest.mark.gpu

M, NR = 16384, 256


@pytest.fixture(scope="module")
def alg():
    from icp_b200 import algorithms
    return algorithms


def second_gpu():
    from icp_b200 import capi
    try:
        c = capi.Context(1)
    except (ValueError, capi.ICPCudaError):
        return None
    return c


def make_pairs(n, seed0=300):
    Fs, Ms = [], []
    for p in range(n):
        F, Mv, _, _ = scene_pair(seed=seed0 + p, deg=1.0 + 0.5 * p, t=(5.0 + p, -3.0, 2.0 * p))
        Fs.append(F); Ms.append(Mv)
    return np.stack(Fs), np.stack(Ms)


def test_multi_on_one_device_matches_oracle(ctx, po, alg):
    n, K = 5, 6
    hF, hM = make_pairs(n)
    mg = alg.ICPMulti(n, M, NR, n_devices=1)
    assert mg.devices() == 1 and mg.pair_range(0) == (0, 0, n)
    T8 = mg.register_host(hF, hM, K)
    T8b = mg.register_host(hF, hM, K)            # cached graphs, same answer
    mg.close()
    for p in range(n):
        ref = po.icp_register(hF[p], hM[p], 128, 128, NR, fixed_iters=K)
        assert_bits_equal(T8[p], ref["T"], f"multi (1 device) pose {p}")
    assert_bits_equal(T8b, T8, "second call")


def test_multi_clamps_devices_to_pairs_and_rejects_bad_arguments(ctx, alg):
    mg = alg.ICPMulti(1, M, NR, n_devices=0)     # all visible GPUs requested, one pair: one device used
    assert mg.devices() == 1
    mg.close()
    with pytest.raises(alg.ICPConfigError):
        alg.ICPMulti(2, M, 0, n_devices=1)
    with pytest.raises(ValueError):
        alg.ICPMulti(2, M, NR, devices=[99])


def test_multi_two_devices_match_oracle(ctx, po, alg):
    c1 = second_gpu()
    if c1 is None:
        pytest.skip("needs 2 GPUs")
    c1.close()
    n, K = 7, 6
    hF, hM = make_pairs(n, seed0=340)
    mg = alg.ICPMulti(n, M, NR, n_devices=2)
    assert mg.devices() == 2
    assert mg.pair_range(0) == (0, 0, 4) and mg.pair_range(1) == (1, 4, 3)
    T8 = mg.register_host(hF, hM, K)
    mg.close()
    for p in range(n):
        ref = po.icp_register(hF[p], hM[p], 128, 128, NR, fixed_iters=K)
        assert_bits_equal(T8[p], ref["T"], f"multi (2 devices) pose {p}")


def test_two_contexts_on_two_devices_in_one_process(ctx, po, alg):
    """entry points that do not set the device launch on whatever device is current.  Interleave stage calls,
    engine runs and copies of two contexts on different GPUs; the caller's current device must not matter."""
    from icp_b200 import capi
    c1 = second_gpu()
    if c1 is None:
        pytest.skip("needs 2 GPUs")
    F, Mv, _, _ = scene_pair(seed=77)
    ref = po.icp_register(F, Mv, 128, 128, NR, fixed_iters=5)
    steps = []
    for c in (ctx, c1, ctx, c1):
        s = alg.ICPStep(c, 1, 1)
        s.init(M, NR, 2e2, 1e-6)
        s.write(capi.MEM_D_IN_F, F); s.write(capi.MEM_D_IN_M, Mv)
        steps.append(s)
    for s in steps:
        s.buildRBC()
    for k in range(5):
        for s in reversed(steps) if k % 2 else steps:
            s.run(1)
    for i, s in enumerate(steps):
        assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], f"engine {i} (device {i % 2})")
    # stage entry points + timers + memset on the second device while device 0 is the process's current device
    T0 = np.array([0.5144, 0.5743, 0.5632, 0.2973, 1.0, -2.0, 3.0, 1.0], np.float32)
    want = po.transform_q(Mv, T0)
    for c in (c1, ctx):
        dM, dT, dO = c.upload(Mv), c.upload(T0), c.alloc(M * 8 * 4)
        c.timer_start()
        capi.check(capi.lib().icp_transform_quaternion(c.h, dM.ptr, dT.ptr, dO.ptr, M))
        assert c.timer_stop() >= 0.0
        assert_bits_equal(capi.read_ptr(c, dO.ptr, np.float32, (M, 8)), want, f"transform on device {c.device}")
    for s in steps:
        s.close()
    c1.close()


def test_mode_switch_between_build_and_run(ctx, po, alg):
    """fused build on frame 1, STAGED build on frame 2, then FUSED iterations -- the fused kernels' acceleration
    tables (representative neighbour rows, temporal bounds, lane order) must follow the new fixed set."""
    from icp_b200 import capi
    F1, M1, _, _ = scene_pair(seed=501)
    F2, M2, _, _ = scene_pair(seed=502, deg=2.0, t=(-12.0, 7.0, 4.0))
    s = alg.ICPStep(ctx, 1, 1)
    s.init(M, NR, 2e2, 1e-6)
    s.set_mode(capi.MODE_FUSED)
    s.write(capi.MEM_D_IN_F, F1); s.write(capi.MEM_D_IN_M, M1)
    s.buildRBC(); s.run(3)
    s.set_mode(capi.MODE_STAGED)
    s.reset()
    s.write(capi.MEM_D_IN_F, F2[::-1].copy()); s.write(capi.MEM_D_IN_M, M2)      # a different fixed set (reversed order: other representatives)
    s.buildRBC()
    s.run(2)                                                                        # staged iterations: k > 0 without a fused lane order
    s.set_mode(capi.MODE_FUSED)
    s.run(4)
    ref = po.icp_register(F2[::-1].copy(), M2, 128, 128, NR, fixed_iters=6, dumps=True)
    assert np.array_equal(s.debug("NN_ID", alg.DIST_ID, M)["id"], ref["nn_id_hist"][5])
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "pose after staged build + staged / fused iterations")
    # new moving set without a new buildRBC (legal: the RBC depends on F only), fused mode throughout
    s.reset()
    s.write(capi.MEM_D_IN_M, M1)
    s.run(3)
    ref = po.icp_register(F2[::-1].copy(), M1, 128, 128, NR, fixed_iters=3)
    assert_bits_equal(s.debug("T", np.float32, 8), ref["T"], "pose after replacing the moving set")
    s.close()
