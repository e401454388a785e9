"""ctypes bindings of the CPU oracle (oracle/libicp_oracle.so) and, when built, of the reference's own
CPU helpers (oracle/_ref/libicp_ref.so).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from icp_b200/ (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC = None
_REF = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def _cpu_fingerprint():
    """ISA of this host (the oracle is compiled -march=native and the .so travels with the repo snapshot)."""
    import hashlib
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return hashlib.sha1(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present). Building is not using."""
    so = os.path.join(_HERE, "libicp_oracle.so")
    stamp = os.path.join(_HERE, ".build_host")
    fp = _cpu_fingerprint()
    try:
        same_host = open(stamp).read().strip() == fp
    except OSError:
        same_host = False
    srcs = [os.path.join(_HERE, "icp_oracle.cpp"), os.path.join(_HERE, "Makefile")]
    if force or not same_host or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        env = dict(os.environ)
        env["CXX"] = "g++"
        subprocess.check_call(["make", "-B", "-C", _HERE, "CXX=g++", "libicp_oracle.so"], env=env, stdout=subprocess.DEVNULL)
        with open(stamp, "w") as fh:
            fh.write(fp + "\n")
    if os.path.isdir("/root/reference") and not os.path.exists(os.path.join(_HERE, "_ref", "libicp_ref.so")):
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++", "ref"], stdout=subprocess.DEVNULL)


class OrcState(C.Structure):
    _fields_ = [("R", C.c_float * 9), ("q", C.c_float * 4), ("t", C.c_float * 3), ("s", C.c_float),
                ("Rk", C.c_float * 9), ("qk", C.c_float * 4), ("tk", C.c_float * 3), ("sk", C.c_float)]


class OrcDumps(C.Structure):
    _fields_ = [("T_hist", C.c_void_p), ("Tk_hist", C.c_void_p), ("nn_id_hist", C.c_void_p),
                ("qperm_hist", C.c_void_p), ("S_hist", C.c_void_p), ("mean_hist", C.c_void_p),
                ("sumw_hist", C.c_void_p), ("e2_hist", C.c_void_p)]


def lib():
    global _ORC
    if _ORC is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libicp_oracle.so"))
        L.orc_get_lms.argtypes = [f32p, f32p]
        L.orc_rgbd_to_pc8d.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, f32p]
        L.orc_rep_grid.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_get_reps.argtypes = [f32p, C.c_uint32, C.c_uint32, C.c_uint32, f32p]
        L.orc_metric_weights.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_nearest_rep.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_float, u32p, C.c_void_p]
        L.orc_counting_sort.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.orc_rbc_construct.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_float, u32p, u32p, u32p, u32p, f32p]
        L.orc_rbc_search.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_float, f32p, u32p, u32p,
                                     u32p, u32p, u32p, u32p, f32p, f32p, f32p, u32p]
        L.orc_nearest_exact.argtypes = [f32p, C.c_uint32, f32p, C.c_uint32, C.c_float, u32p, f32p]
        L.orc_transform_q.argtypes = [f32p, C.c_uint32, f32p, f32p]
        L.orc_transform_m.argtypes = [f32p, C.c_uint32, f32p, f32p]
        L.orc_weights.argtypes = [f32p, C.c_uint32, f32p, C.POINTER(C.c_double)]
        L.orc_mean.argtypes = [f32p, f32p, C.c_uint32, f32p]
        L.orc_mean_weighted.argtypes = [f32p, f32p, f32p, C.c_double, C.c_uint32, f32p]
        L.orc_devs.argtypes = [f32p, f32p, f32p, C.c_uint32, f32p, f32p]
        L.orc_sij.argtypes = [f32p, f32p, C.c_void_p, C.c_uint32, C.c_float, f32p]
        L.orc_reduce_sum_f.argtypes = [f32p, C.c_uint32, C.c_uint32, f32p]
        L.orc_reduce_min_f.argtypes = [f32p, C.c_uint32, C.c_uint32, f32p]
        L.orc_reduce_max_ui.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p]
        L.orc_scan_i.argtypes = [i32p, C.c_uint32, C.c_uint32, C.c_int, i32p]
        L.orc_power_method.argtypes = [f32p, f32p, f32p]
        L.orc_power_method.restype = C.c_int
        L.orc_svd_solve.argtypes = [f32p, f32p, f32p, f32p]
        L.orc_state_init.argtypes = [C.POINTER(OrcState)]
        L.orc_accumulate.argtypes = [C.POINTER(OrcState), f32p, C.c_void_p, f32p]
        L.orc_converged.argtypes = [f32p, f32p, C.c_double, C.c_double]
        L.orc_converged.restype = C.c_int
        L.orc_pose_matrix.argtypes = [C.POINTER(OrcState), f32p]
        L.orc_icp_register.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                       C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_double, C.c_double,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_icp_register.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        L.orc_hw_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _ORC = L
    return _ORC


def ref_lib():
    """The reference's own CPU helpers (None when oracle/_ref was not built: no /root/reference)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libicp_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        vp = C.c_void_p
        L.ref_ICPLMs.argtypes = [f32p, f32p]
        L.ref_ICPReps.argtypes = [f32p, f32p, C.c_uint32]
        L.ref_ICPWeights.argtypes = [vp, f32p, C.POINTER(C.c_double), C.c_uint32]
        L.ref_ICPMean.argtypes = [f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPMeanWeighted.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPDevs.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPS.argtypes = [f32p, f32p, f32p, C.c_uint32, C.c_float]
        L.ref_ICPSw.argtypes = [f32p, f32p, f32p, f32p, C.c_uint32, C.c_float]
        L.ref_ICPTransformQ.argtypes = [f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPTransformQ2.argtypes = [f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPTransformM.argtypes = [f32p, f32p, f32p, C.c_uint32]
        L.ref_ICPPowerMethod.argtypes = [f32p, f32p, f32p]
        L.ref_ReduceSum.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32]
        L.ref_ReduceMin.argtypes = [f32p, f32p, C.c_uint32, C.c_uint32]
        L.ref_ReduceMaxU.argtypes = [u32p, u32p, C.c_uint32, C.c_uint32]
        L.ref_InScan.argtypes = [i32p, i32p, C.c_uint32, C.c_uint32]
        L.ref_ExScan.argtypes = [i32p, i32p, C.c_uint32, C.c_uint32]
        _REF = L
    return _REF


# ---------------------------------------------------------------------------------------------------------
# numpy-level wrappers
# ---------------------------------------------------------------------------------------------------------
def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def get_lms(cloud):
    out = np.empty((16384, 8), np.float32)
    lib().orc_get_lms(_f(cloud).reshape(-1), out.reshape(-1))
    return out


def rgbd_to_pc8d(depth, rgb, f=595.0):
    """depth (H, W) uint16, rgb (H, W, 3) uint8 -> (H*W, 8) float32 (kinect_frame_grabber.cpp:246-263)"""
    depth = np.ascontiguousarray(depth, np.uint16)
    rgb = np.ascontiguousarray(rgb, np.uint8)
    H, W = depth.shape
    out = np.empty((H * W, 8), np.float32)
    lib().orc_rgbd_to_pc8d(depth.ctypes.data, rgb.ctypes.data, W, H, f, out.reshape(-1))
    return out


def rep_grid(nr):
    a, b = C.c_uint32(), C.c_uint32()
    lib().orc_rep_grid(nr, C.byref(a), C.byref(b))
    return a.value, b.value


def get_reps(lms, W, H, nr):
    out = np.empty((nr, 8), np.float32)
    lib().orc_get_reps(_f(lms).reshape(-1), W, H, nr, out.reshape(-1))
    return out


def metric_weights(a):
    fg, fp = C.c_float(), C.c_float()
    lib().orc_metric_weights(a, C.byref(fg), C.byref(fp))
    return fg.value, fp.value


def nearest_rep(X, R, a):
    X = _f(X); R = _f(R)
    ids = np.empty(len(X), np.uint32)
    d = np.empty(len(X), np.float32)
    lib().orc_nearest_rep(X.reshape(-1), len(X), R.reshape(-1), len(R), a, ids, d.ctypes.data)
    return ids, d


def counting_sort(key, nr):
    key = np.ascontiguousarray(key, np.uint32)
    N = np.empty(nr, np.uint32); O = np.empty(nr, np.uint32); perm = np.empty(len(key), np.uint32)
    lib().orc_counting_sort(key, len(key), nr, N, O, perm)
    return N, O, perm


def rbc_construct(X, R, a):
    X = _f(X); R = _f(R)
    n, nr = len(X), len(R)
    rep_id = np.empty(n, np.uint32); N = np.empty(nr, np.uint32); O = np.empty(nr, np.uint32)
    perm = np.empty(n, np.uint32); Xp = np.empty((n, 8), np.float32)
    lib().orc_rbc_construct(X.reshape(-1), n, R.reshape(-1), nr, a, rep_id, N, O, perm, Xp.reshape(-1))
    return dict(rep_id=rep_id, N=N, O=O, perm=perm, Xp=Xp)


def rbc_search(Q, R, a, Xp, O, N):
    Q = _f(Q); R = _f(R); Xp = _f(Xp)
    m, nr = len(Q), len(R)
    q_rep = np.empty(m, np.uint32); Nq = np.empty(nr, np.uint32); Oq = np.empty(nr, np.uint32)
    qperm = np.empty(m, np.uint32); Qp = np.empty((m, 8), np.float32); NN = np.empty((m, 8), np.float32)
    nn_dist = np.empty(m, np.float32); nn_id = np.empty(m, np.uint32)
    lib().orc_rbc_search(Q.reshape(-1), m, R.reshape(-1), nr, a, Xp.reshape(-1),
                         np.ascontiguousarray(O, np.uint32), np.ascontiguousarray(N, np.uint32),
                         q_rep, Nq, Oq, qperm, Qp.reshape(-1), NN.reshape(-1), nn_dist, nn_id)
    return dict(q_rep=q_rep, Nq=Nq, Oq=Oq, qperm=qperm, Qp=Qp, NN=NN, nn_dist=nn_dist, nn_id=nn_id)


def nearest_exact(Q, Xp, a):
    """Exact nearest neighbour over the whole list-ordered database (brute force; lowest list position among ties)."""
    Q, Xp = _f(Q), _f(Xp)
    m, n = len(Q), len(Xp)
    nn_id = np.zeros(m, np.uint32)
    nn_dist = np.zeros(m, np.float32)
    lib().orc_nearest_exact(Q, m, Xp, n, a, nn_id, nn_dist)
    return dict(nn_id=nn_id, nn_dist=nn_dist)


def transform_q(M, T):
    M = _f(M); out = np.empty_like(M)
    lib().orc_transform_q(M.reshape(-1), len(M), _f(T).reshape(-1), out.reshape(-1))
    return out


def transform_m(M, T16):
    M = _f(M); out = np.empty_like(M)
    lib().orc_transform_m(M.reshape(-1), len(M), _f(T16).reshape(-1), out.reshape(-1))
    return out


def weights(dist):
    dist = _f(dist)
    W = np.empty_like(dist); s = C.c_double()
    lib().orc_weights(dist, len(dist), W, C.byref(s))
    return W, s.value


def mean(F, M):
    out = np.empty(8, np.float32)
    lib().orc_mean(_f(F).reshape(-1), _f(M).reshape(-1), len(F), out)
    return out


def mean_weighted(F, M, W, sum_w):
    out = np.empty(8, np.float32)
    lib().orc_mean_weighted(_f(F).reshape(-1), _f(M).reshape(-1), _f(W), sum_w, len(F), out)
    return out


def devs(F, M, mean8):
    n = len(F)
    DF = np.empty((n, 4), np.float32); DM = np.empty((n, 4), np.float32)
    lib().orc_devs(_f(F).reshape(-1), _f(M).reshape(-1), _f(mean8), n, DF.reshape(-1), DM.reshape(-1))
    return DF, DM


def sij(DM, DF, W, c):
    S = np.empty(11, np.float32)
    Wp = None if W is None else _f(W).ctypes.data
    if W is not None:
        W = _f(W); Wp = W.ctypes.data
    lib().orc_sij(_f(DM).reshape(-1), _f(DF).reshape(-1), Wp, len(DM), c, S)
    return S


def power_method(S, means):
    Tk = np.empty(8, np.float32)
    it = lib().orc_power_method(_f(S), _f(means), Tk)
    return Tk, it


def svd_solve(S, means):
    Tk = np.empty(8, np.float32); Rk = np.empty(9, np.float32)
    lib().orc_svd_solve(_f(S), _f(means), Tk, Rk)
    return Tk, Rk.reshape(3, 3)


def reduce_sum_f(a):
    a = _f(a); out = np.empty(a.shape[0], np.float32)
    lib().orc_reduce_sum_f(a.reshape(-1), a.shape[1], a.shape[0], out)
    return out


def reduce_min_f(a):
    a = _f(a); out = np.empty(a.shape[0], np.float32)
    lib().orc_reduce_min_f(a.reshape(-1), a.shape[1], a.shape[0], out)
    return out


def reduce_max_ui(a):
    a = np.ascontiguousarray(a, np.uint32); out = np.empty(a.shape[0], np.uint32)
    lib().orc_reduce_max_ui(a.reshape(-1), a.shape[1], a.shape[0], out)
    return out


def scan_i(a, inclusive):
    a = np.ascontiguousarray(a, np.int32); out = np.empty_like(a)
    lib().orc_scan_i(a.reshape(-1), a.shape[1], a.shape[0], 1 if inclusive else 0, out.reshape(-1))
    return out


def icp_register(F, M, W, H, nr, a=2e2, c=1e-6, rot="power", weighted=True, fixed_iters=0,
                 max_iterations=40, angle_thr=0.001, trans_thr=0.01, T0=None, dumps=False):
    """Full registration (ICPStep::buildRBC + ICP::run).  Returns dict(k, T, T16, [histories])."""
    F = _f(F); M = _f(M)
    m = len(F)
    T = np.empty(8, np.float32); T16 = np.empty(16, np.float32)
    K = fixed_iters if fixed_iters > 0 else max_iterations
    d = None; hist = {}
    if dumps:
        hist = dict(T_hist=np.zeros((K, 8), np.float32), Tk_hist=np.zeros((K, 8), np.float32),
                    nn_id_hist=np.zeros((K, m), np.uint32), qperm_hist=np.zeros((K, m), np.uint32),
                    S_hist=np.zeros((K, 11), np.float32), mean_hist=np.zeros((K, 8), np.float32),
                    sumw_hist=np.zeros(K, np.float64), e2_hist=np.zeros(K, np.uint64))
        d = OrcDumps(*[hist[k].ctypes.data for k in
                       ("T_hist", "Tk_hist", "nn_id_hist", "qperm_hist", "S_hist", "mean_hist", "sumw_hist", "e2_hist")])
    T0p = None
    if T0 is not None:
        T0 = _f(T0); T0p = T0.ctypes.data
    k = lib().orc_icp_register(F.reshape(-1), M.reshape(-1), m, W, H, nr, a, c,
                               1 if rot == "power" else 0, 1 if weighted else 0, fixed_iters, max_iterations,
                               angle_thr, trans_thr, T0p, T.ctypes.data, T16.ctypes.data,
                               C.byref(d) if d is not None else None)
    out = dict(k=k, T=T, T16=T16.reshape(4, 4))
    for kk, v in hist.items():
        out[kk] = v[:k]
    return out


def set_threads(n):
    lib().orc_set_num_threads(int(n))


def hw_threads():
    return lib().orc_hw_threads()
