"""RGB-D -> pc8d front end: the conversion the reference's frame grabber applies before it writes
kg_pc8d_*.bin (/root/reference/src/kinect_frame_grabber.cpp:246-263).  Oracle vs a literal numpy restatement on CPU;
CUDA kernel vs oracle, bit-exact, on the GPU (full frame, ragged sizes, invalid depth)."""
import numpy as np
import pytest

from util import assert_bits_equal


def frames(W, H, seed):
    rng = np.random.default_rng(seed)
    depth = rng.integers(0, 10001, (H, W), dtype=np.uint16)          # the reference tests' ushort U[0, 10000]
    depth[rng.uniform(size=(H, W)) < 0.08] = 0                        # invalid pixels
    rgb = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    return depth, rgb


def numpy_restatement(depth, rgb, f):
    H, W = depth.shape
    x = np.arange(W, dtype=np.float32)[None, :].repeat(H, 0)
    y = np.arange(H, dtype=np.float32)[:, None].repeat(W, 1)
    d = depth.astype(np.float32)
    out = np.ones((H, W, 8), np.float32)
    out[..., 0] = (x - np.float32((W - 1) / 2.0)) * d / np.float32(f)
    out[..., 1] = (y - np.float32((H - 1) / 2.0)) * d / np.float32(f)
    out[..., 2] = d
    out[..., 4:7] = rgb.astype(np.float32) / np.float32(255.0)
    return out.reshape(-1, 8)


@pytest.mark.parametrize("W,H", [(640, 480), (37, 5), (1, 1)])
def test_oracle_is_the_grabber_formula(po, W, H):
    depth, rgb = frames(W, H, 3)
    got = po.rgbd_to_pc8d(depth, rgb, 595.0)
    assert_bits_equal(got, numpy_restatement(depth, rgb, 595.0), "pc8d")
    inv = depth.reshape(-1) == 0
    assert np.all(got[inv, :3] == 0) and np.all(got[:, 3] == 1) and np.all(got[:, 7] == 1)


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,f", [(640, 480, 595.0), (641, 479, 520.5), (33, 7, 595.0), (1, 1, 1.0), (2048, 1536, 1050.0)])
def test_kernel_matches_oracle(ctx, po, W, H, f):
    from icp_b200 import algorithms as alg
    depth, rgb = frames(W, H, 11)
    k = alg.RGBDTo8D(ctx)
    k.init(W, H, f)
    k.write("D_IN_D", depth)
    k.write("D_IN_RGB", rgb)
    k.run()
    assert_bits_equal(k.read(), po.rgbd_to_pc8d(depth, rgb, f), f"pc8d {W}x{H}")


@pytest.mark.gpu
def test_frontend_feeds_the_landmark_sampler(ctx, po):
    """depth + rgb -> pc8d -> ICPLMs: the path of a raw Kinect frame into the registration."""
    from icp_b200 import algorithms as alg
    depth, rgb = frames(640, 480, 5)
    k = alg.RGBDTo8D(ctx)
    k.init()
    k.write("D_IN_D", depth); k.write("D_IN_RGB", rgb)
    lms = alg.ICPLMs(ctx)
    lms.set("D_IN", k.get("D_OUT"))           # `lms.get (Memory::D_IN) = k.get (Memory::D_OUT)` before init (), as in the reference
    lms.init()
    k.run()
    lms.run()
    assert_bits_equal(lms.read(), po.get_lms(po.rgbd_to_pc8d(depth, rgb)), "landmarks of the converted frame")


@pytest.mark.gpu
def test_config_errors(ctx):
    from icp_b200 import algorithms as alg
    k = alg.RGBDTo8D(ctx)
    k.init(8, 8, 0.0)
    with pytest.raises(alg.ICPConfigError):
        k.run()
