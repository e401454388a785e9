"""Seeded synthetic pc8d scenes (stand-ins for the reference's bundled data, which is absent:
/root/reference/.MISSING_LARGE_BLOBS) and the `.bin` loader.

pc8d layout = 640x480 points of 8 x f32 `[x, y, z, 1, r, g, b, 1]`, xyz in mm through the grabber's pinhole
(f = 595, centre (319.5, 239.5): /root/reference/src/kinect_frame_grabber.cpp:253-261), rgb in [0,1],
invalid pixels have xyz = 0.  File format: raw little-endian f32, 640*480*8 values
(/root/reference/examples/registration.cpp:285-337).

Scenes follow SURVEY.md section 8(d): S-room (config 1), S-wall (config 2), known-transform landmark
pairs (config 3 / 5).  Pure numpy; deterministic for a given seed.
"""
import os

import numpy as np

W, H = 640, 480
FOCAL = 595.0
CX, CY = 319.5, 239.5


def load_pc8d(path):
    """Read a kg_pc8d_*.bin cloud -> (307200, 8) float32."""
    a = np.fromfile(path, dtype="<f4")
    if a.size != W * H * 8:
        raise ValueError(f"{path}: expected {W * H * 8} floats, found {a.size}")
    return a.reshape(W * H, 8)


def save_pc8d(path, cloud):
    np.ascontiguousarray(cloud, dtype="<f4").reshape(-1).tofile(path)


def find_bundled_pair(name="kg_pc8d", data_dir=None):
    """Return (path1, path2) of a bundled pair if the blobs were dropped into data/, else None."""
    data_dir = data_dir or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")
    p1, p2 = os.path.join(data_dir, f"{name}_1.bin"), os.path.join(data_dir, f"{name}_2.bin")
    return (p1, p2) if os.path.exists(p1) and os.path.exists(p2) else None


def axis_angle(axis, deg):
    """Rotation matrix (float64) from axis / angle in degrees."""
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    th = np.deg2rad(deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def rot_to_quat(R):
    """[x, y, z, w] from a rotation matrix (float64, Shoemake)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        return np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    i = int(np.argmax(np.diag(R)))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
    q = np.zeros(4)
    q[i] = 0.25 * s
    q[3] = (R[k, j] - R[j, k]) / s
    q[j] = (R[j, i] + R[i, j]) / s
    q[k] = (R[k, i] + R[i, k]) / s
    return q


def _rays():
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    return np.stack([(u - CX) / FOCAL, (v - CY) / FOCAL, np.ones_like(u)], -1).reshape(-1, 3)


def _texture(P, sid, seed):
    """Procedural rgb in [0,1] per surface id: 3 sinusoids + checker."""
    rng = np.random.default_rng(seed)
    rgb = np.zeros((len(P), 3))
    for s in np.unique(sid):
        sel = sid == s
        if s < 0:
            continue
        rs = np.random.default_rng(seed * 131 + int(s))
        p = P[sel]
        col = np.zeros((sel.sum(), 3))
        for c in range(3):
            acc = 0.5 * np.ones(sel.sum())
            for _ in range(3):
                k = rs.uniform(-1, 1, 3) * rs.uniform(0.004, 0.03)
                acc += rs.uniform(0.08, 0.16) * np.sin(p @ k + rs.uniform(0, 6.28))
            col[:, c] = acc
        chk = ((np.floor(p[:, 0] / 150.0) + np.floor(p[:, 1] / 150.0) + np.floor(p[:, 2] / 150.0)) % 2)
        col += (chk[:, None] - 0.5) * 0.18
        rgb[sel] = col
    del rng
    return np.clip(rgb, 0.0, 1.0)


def _render(surfaces, Rc, tc, seed, tex_seed, invalid_frac=0.08, block_tex=None):
    """Ray-cast `surfaces` from a camera with pose (Rc, tc) (camera->world).  Returns (307200, 8) float32."""
    rng = np.random.default_rng(seed)
    d_cam = _rays()
    d_w = d_cam @ Rc.T
    o = tc
    lam = np.full(len(d_cam), np.inf)
    sid = np.full(len(d_cam), -1, np.int64)
    for i, s in enumerate(surfaces):
        if s[0] == "plane":           # n . x = c
            n, c = np.asarray(s[1], np.float64), s[2]
            den = d_w @ n
            with np.errstate(divide="ignore", invalid="ignore"):
                l = (c - o @ n) / den
            ok = (np.abs(den) > 1e-9) & (l > 1.0)
        elif s[0] == "box":
            lo, hi = np.asarray(s[1], np.float64), np.asarray(s[2], np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                t1 = (lo - o) / d_w
                t2 = (hi - o) / d_w
            tmin = np.max(np.minimum(t1, t2), axis=1)
            tmax = np.min(np.maximum(t1, t2), axis=1)
            l = tmin
            ok = (tmax >= tmin) & (tmin > 1.0)
        else:                          # sphere
            cen, r = np.asarray(s[1], np.float64), s[2]
            oc = o - cen
            a = np.sum(d_w * d_w, axis=1)
            b = 2 * (d_w @ oc)
            cc = oc @ oc - r * r
            disc = b * b - 4 * a * cc
            ok = disc > 0
            l = (-b - np.sqrt(np.where(ok, disc, 0))) / (2 * a)
            ok &= l > 1.0
        upd = ok & (l < lam)
        lam[upd] = l[upd]
        sid[upd] = i
    hit = np.isfinite(lam)
    z = np.where(hit, lam, 0.0)
    Pw = o + d_w * z[:, None]
    if block_tex is None:
        rgb = _texture(Pw, sid, tex_seed)
    else:
        rgb = block_tex(Pw)
    # depth range of the sensor
    hit &= (z > 400.0) & (z < 6000.0)
    # depth-edge shadows: pixels right of a large depth jump
    zi = z.reshape(H, W)
    jump = np.abs(np.diff(zi, axis=1, prepend=zi[:, :1])) > 120.0
    shadow = jump.copy()
    for k in range(1, 4):
        shadow[:, k:] |= jump[:, :-k]
    invalid = ~hit | shadow.reshape(-1)
    # random invalid blobs up to the requested share
    vv, uu = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    tries = 0
    while invalid.mean() < invalid_frac and tries < 400:
        cu, cv, rr = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(6, 28)
        invalid |= (((uu - cu) ** 2 + (vv - cv) ** 2) < rr * rr).reshape(-1)
        tries += 1
    # sensor noise
    zn = z + rng.normal(0, 1.0, len(z)) * 1.5 * (z / 1000.0) ** 2
    rgb = np.clip(rgb + rng.normal(0, 0.01, rgb.shape), 0.0, 1.0)
    out = np.zeros((W * H, 8), np.float32)
    out[:, 0] = d_cam[:, 0] * zn
    out[:, 1] = d_cam[:, 1] * zn
    out[:, 2] = zn
    out[invalid, 0:3] = 0.0
    out[:, 3] = 1.0
    out[:, 4:7] = rgb
    out[:, 7] = 1.0
    return out


ROOM = [
    ("plane", (0, 1, 0), 1000.0),                       # floor (y down)
    ("plane", (0, 0, 1), 3000.0),                       # back wall
    ("plane", (1, 0, 0), -2200.0), ("plane", (1, 0, 0), 2300.0),   # side walls
    ("box", (-900, 400, 1500), (-300, 1000, 2100)),
    ("box", (250, 100, 2000), (900, 1000, 2600)),
    ("box", (-250, 650, 1100), (150, 1000, 1500)),
    ("sphere", (450, -150, 1700), 330.0),
]

ROOM_GT = dict(axis=(0.2, 1.0, 0.1), deg=3.0, t=(25.0, -10.0, 15.0))


def room_pair(seed=1001):
    """Config 1 stand-in (S-room).  Returns (cloud1, cloud2, R_gt, t_gt): p_1 = R_gt p_2 + t_gt."""
    R = axis_angle(ROOM_GT["axis"], ROOM_GT["deg"])
    t = np.asarray(ROOM_GT["t"], np.float64)
    c1 = _render(ROOM, np.eye(3), np.zeros(3), seed, tex_seed=77)
    c2 = _render(ROOM, R, t, seed + 1, tex_seed=77)
    return c1, c2, R, t


def wall_pair(seed=2001):
    """Config 2 stand-in (S-wall): one slanted plane, block texture; in-plane motion."""
    rs = np.random.default_rng(seed)
    blocks = rs.uniform(0, 1, (64, 64, 3))
    n = np.array([-0.02, 0.0, 1.0]); n /= np.linalg.norm(n)
    surf = [("plane", tuple(n), 1500.0 * n[2])]

    def tex(P):
        bu = np.clip(((P[:, 0] + 2000) / 62.5).astype(np.int64), 0, 63)
        bv = np.clip(((P[:, 1] + 2000) / 62.5).astype(np.int64), 0, 63)
        shade = 0.15 * np.sin(P[:, 0] * 0.004)[:, None] + 0.1 * np.cos(P[:, 1] * 0.003)[:, None]
        return np.clip(0.7 * blocks[bv, bu] + 0.15 + shade, 0, 1)

    R = axis_angle((0, 0, 1), 1.0)
    t = np.array([30.0, 12.0, 0.0])
    c1 = _render(surf, np.eye(3), np.zeros(3), seed, 0, invalid_frac=0.03, block_tex=tex)
    c2 = _render(surf, R, t, seed + 1, 0, invalid_frac=0.03, block_tex=tex)
    return c1, c2, R, t


def landmarks_np(cloud):
    """numpy statement of the 128x128 landmark sampling (row 49+3*gy, column 65+4*lx) -- data prep for
    synthetic pairs only (the product path uses the CUDA kernel)."""
    g = cloud.reshape(H, W, 8)
    return np.ascontiguousarray(g[49:49 + 3 * 128:3, 65:65 + 4 * 128:4].reshape(-1, 8))


_BASE_CACHE = {}


def base_landmarks(seed=3001):
    if seed not in _BASE_CACHE:
        _BASE_CACHE[seed] = landmarks_np(_render(ROOM, np.eye(3), np.zeros(3), seed, tex_seed=77))
    return _BASE_CACHE[seed]


def grid_cloud(Wg, Hg, seed=3001):
    """A Wg x Hg grid sampled from the S-room frame (configs 4: 256x256, 640x480)."""
    g = _render(ROOM, np.eye(3), np.zeros(3), seed, tex_seed=77).reshape(H, W, 8)
    if (Wg, Hg) == (W, H):
        return np.ascontiguousarray(g.reshape(-1, 8))
    sx, sy = W // Wg, H // Hg
    if sx >= 1 and sy >= 1:
        ox, oy = (W - sx * Wg) // 2, (H - sy * Hg) // 2
        return np.ascontiguousarray(g[oy:oy + sy * Hg:sy, ox:ox + sx * Wg:sx].reshape(-1, 8))
    raise ValueError("grid larger than the frame")


def known_transform_pair(seed=3001, deg=4.0, t=(20.0, -15.0, 10.0), axis=None, F=None,
                         xyz_sigma=1.0, rgb_sigma=0.005, outliers=0.05):
    """Config 3: F = landmarks; M = T_gt^-1(F) + noise (+ outliers).  Returns (F, M, R_gt, t_gt)."""
    rng = np.random.default_rng(seed)
    if F is None:
        F = base_landmarks(3001)
    if axis is None:
        axis = (0.5144, 0.5743, 0.5632)
    R = axis_angle(axis, deg)
    t = np.asarray(t, np.float64)
    M = F.astype(np.float64).copy()
    valid = np.any(F[:, :3] != 0, axis=1)
    M[:, :3] = (M[:, :3] - t) @ R            # R^T (p - t) written row-wise
    M[:, :3] += rng.normal(0, xyz_sigma, (len(M), 3))
    M[:, 4:7] = np.clip(M[:, 4:7] + rng.normal(0, rgb_sigma, (len(M), 3)), 0, 1)
    M[~valid, :3] = 0.0
    n_out = int(outliers * len(M))
    if n_out:
        idx = rng.choice(len(M), n_out, replace=False)
        lo, hi = F[valid, :3].min(0), F[valid, :3].max(0)
        M[idx, :3] = rng.uniform(lo, hi, (n_out, 3))
        M[idx, 4:7] = rng.uniform(0, 1, (n_out, 3))
    return F.astype(np.float32), M.astype(np.float32), R, t


def batch_pair(i, F=None):
    """Config 5, pair i: random axis, angle U[0.5,5] deg, |t| U[5,50] mm, seed 5000+i."""
    rng = np.random.default_rng(5000 + i)
    axis = rng.normal(size=3)
    deg = rng.uniform(0.5, 5.0)
    tdir = rng.normal(size=3)
    t = tdir / np.linalg.norm(tdir) * rng.uniform(5.0, 50.0)
    return known_transform_pair(seed=5000 + i, deg=deg, t=t, axis=axis, F=F)
