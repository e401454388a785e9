"""Python mirror of the reference's algorithm-class API (cl_algo::ICP::*, /root/reference/include/ICP/algorithms.hpp)
on top of the C ABI.  Same class names, same init / write / read / run rhythm, same error behaviour
(configuration errors raise ICPConfigError where the reference prints `Error[<Class>]` and exits).

The C++ drop-in (include/ICP/algorithms.hpp) is the product boundary for C++ callers; this module is what the
pytest suite and bench.py drive.  No oracle, no CPU fallback in here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (Context, DeviceBuffer, ICPConfigError, ICPCudaError, check, lib,
                   ROT_EIGEN, ROT_POWER_METHOD, W_REGULAR, W_WEIGHTED, MODE_STAGED, MODE_FUSED)

DIST_ID = np.dtype([("dist", np.float32), ("id", np.uint32)])


class _Stage:
    """Uniform shape of the reference stage classes: ctor(env/info) . init(sizes) . write . run . read."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.buf = {}

    def _alloc(self, name, nbytes):
        if name not in self.buf or self.buf[name] is None:          # init() only creates what is still null
            self.buf[name] = self.ctx.alloc(nbytes)
        return self.buf[name]

    def get(self, name):
        return self.buf.get(name)

    def set(self, name, devbuf):
        """`obj.get (Memory::X) = buffer` before init(): share a device buffer between stages."""
        self.buf[name] = devbuf

    def write(self, name, arr):
        self.buf[name].write(arr)

    def sync(self):
        self.ctx.sync()


class ICPLMs(_Stage):
    """algorithms.hpp:312-374"""

    def init(self):
        self._alloc("D_IN", 640 * 480 * 32)
        self._alloc("D_OUT", 16384 * 32)

    def run(self):
        check(lib().icp_get_lms(self.ctx.h, self.buf["D_IN"].ptr, self.buf["D_OUT"].ptr))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, (16384, 8))


class RGBDTo8D(_Stage):
    """RGB-D frame -> pc8d cloud (the reference grabber's conversion, src/kinect_frame_grabber.cpp:222, :246-263)."""

    def init(self, W=640, H=480, f=595.0):
        self.W, self.H, self.f = W, H, f
        self._alloc("D_IN_D", W * H * 2)
        self._alloc("D_IN_RGB", W * H * 3)
        self._alloc("D_OUT", W * H * 32)

    def run(self):
        check(lib().icp_rgbd_to_pc8d(self.ctx.h, self.buf["D_IN_D"].ptr, self.buf["D_IN_RGB"].ptr, self.W, self.H, self.f,
                                     self.buf["D_OUT"].ptr))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, (self.W * self.H, 8))


class ICPReps(_Stage):
    """algorithms.hpp:397-459 (W x H landmark grid generalisation, default 128 x 128)"""

    def init(self, nr, W=128, H=128):
        self.nr, self.W, self.H = nr, W, H
        self._alloc("D_IN", W * H * 32)
        self._alloc("D_OUT", max(nr, 1) * 32)

    def run(self):
        check(lib().icp_get_reps(self.ctx.h, self.buf["D_IN"].ptr, self.W, self.H, self.nr, self.buf["D_OUT"].ptr))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, (self.nr, 8))


class ICPTransform(_Stage):
    """algorithms.hpp:1239-1424; config 'QUATERNION' | 'MATRIX'"""

    def __init__(self, ctx, config="QUATERNION"):
        super().__init__(ctx)
        self.config = config

    def init(self, m):
        self.m = m
        self._alloc("D_IN_M", max(m, 1) * 32)
        self._alloc("D_IN_T", 32 if self.config == "QUATERNION" else 64)
        self._alloc("D_OUT", max(m, 1) * 32)

    def run(self):
        fn = lib().icp_transform_quaternion if self.config == "QUATERNION" else lib().icp_transform_matrix
        check(fn(self.ctx.h, self.buf["D_IN_M"].ptr, self.buf["D_IN_T"].ptr, self.buf["D_OUT"].ptr, self.m))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, (self.m, 8))


class RBCConstruct(_Stage):
    """RBC::RBCConstruct<KINECT_R,GENERIC> as wired at algorithms.cpp:4503-4508"""

    def init(self, n, nr, alpha):
        self.n, self.nr, self.alpha = n, nr, alpha
        self._alloc("D_IN_X", max(n, 1) * 32)
        self._alloc("D_IN_R", max(nr, 1) * 32)
        self._alloc("D_OUT_ID", max(n, 1) * 4)
        self._alloc("D_OUT_N", max(nr, 1) * 4)
        self._alloc("D_OUT_O", max(nr, 1) * 4)
        self._alloc("D_OUT_PERM", max(n, 1) * 4)
        self._alloc("D_OUT_X_P", max(n, 1) * 32)

    def run(self):
        b = self.buf
        check(lib().icp_rbc_construct(self.ctx.h, b["D_IN_X"].ptr, self.n, b["D_IN_R"].ptr, self.nr, self.alpha,
                                      b["D_OUT_ID"].ptr, b["D_OUT_N"].ptr, b["D_OUT_O"].ptr, b["D_OUT_PERM"].ptr, b["D_OUT_X_P"].ptr))

    def read(self):
        b = self.buf
        return dict(rep_id=b["D_OUT_ID"].read(np.uint32, self.n), N=b["D_OUT_N"].read(np.uint32, self.nr),
                    O=b["D_OUT_O"].read(np.uint32, self.nr), perm=b["D_OUT_PERM"].read(np.uint32, self.n),
                    Xp=b["D_OUT_X_P"].read(np.float32, (self.n, 8)))


class RBCSearch(_Stage):
    """RBC::RBCSearch<KINECT_R,GENERIC,KINECT> as wired at algorithms.cpp:4520-4536"""

    def init(self, m, nr, alpha):
        self.m, self.nr, self.alpha = m, nr, alpha
        for name, nb in (("D_IN_Q", m * 32), ("D_IN_R", nr * 32), ("D_IN_X_P", m * 32), ("D_IN_O", nr * 4), ("D_IN_N", nr * 4),
                         ("D_OUT_Q_P", m * 32), ("D_OUT_NN", m * 32), ("D_OUT_NN_ID", m * 8), ("D_OUT_Q_REP", m * 4),
                         ("D_OUT_Q_PERM", m * 4), ("D_OUT_NQ", nr * 4), ("D_OUT_OQ", nr * 4)):
            self._alloc(name, max(nb, 16))

    def run(self):
        b = self.buf
        check(lib().icp_rbc_search(self.ctx.h, b["D_IN_Q"].ptr, self.m, b["D_IN_R"].ptr, self.nr, self.alpha, b["D_IN_X_P"].ptr,
                                   b["D_IN_O"].ptr, b["D_IN_N"].ptr, b["D_OUT_Q_P"].ptr, b["D_OUT_NN"].ptr, b["D_OUT_NN_ID"].ptr,
                                   b["D_OUT_Q_REP"].ptr, b["D_OUT_Q_PERM"].ptr, b["D_OUT_NQ"].ptr, b["D_OUT_OQ"].ptr))

    def read(self):
        b = self.buf
        nnid = b["D_OUT_NN_ID"].read(DIST_ID, self.m)
        return dict(Qp=b["D_OUT_Q_P"].read(np.float32, (self.m, 8)), NN=b["D_OUT_NN"].read(np.float32, (self.m, 8)),
                    nn_dist=nnid["dist"].copy(), nn_id=nnid["id"].copy(), q_rep=b["D_OUT_Q_REP"].read(np.uint32, self.m),
                    qperm=b["D_OUT_Q_PERM"].read(np.uint32, self.m), Nq=b["D_OUT_NQ"].read(np.uint32, self.nr),
                    Oq=b["D_OUT_OQ"].read(np.uint32, self.nr))


class RBCSearchExact(_Stage):
    """Exact nearest neighbour over the random ball cover (icp_rbc_search_exact): same database buffers as
    RBCSearch (n database points, m queries, n may differ from m: frame-to-model), results in ORIGINAL query order and
    identical to a brute-force scan of X_p."""

    def init(self, m, n, nr, alpha):
        self.m, self.n, self.nr, self.alpha = m, n, nr, alpha
        for name, nb in (("D_IN_Q", m * 32), ("D_IN_R", nr * 32), ("D_IN_X_P", n * 32), ("D_IN_O", nr * 4), ("D_IN_N", nr * 4),
                         ("D_OUT_NN", m * 32), ("D_OUT_NN_ID", m * 8), ("D_EVALS", 8)):
            self._alloc(name, max(nb, 16))

    def run(self):
        b = self.buf
        check(lib().icp_memset(self.ctx.h, b["D_EVALS"].ptr, 0, 8))
        check(lib().icp_rbc_search_exact(self.ctx.h, b["D_IN_Q"].ptr, self.m, b["D_IN_R"].ptr, self.nr, self.alpha, b["D_IN_X_P"].ptr,
                                         b["D_IN_O"].ptr, b["D_IN_N"].ptr, b["D_OUT_NN_ID"].ptr, b["D_OUT_NN"].ptr, b["D_EVALS"].ptr))

    def read(self):
        b = self.buf
        nnid = b["D_OUT_NN_ID"].read(DIST_ID, self.m)
        return dict(NN=b["D_OUT_NN"].read(np.float32, (self.m, 8)), nn_dist=nnid["dist"].copy(), nn_id=nnid["id"].copy(),
                    evals=int(b["D_EVALS"].read(np.uint64, 1)[0]))


class ICPWeights(_Stage):
    """algorithms.hpp:485-572"""

    def init(self, n):
        self.n = n
        self._alloc("D_IN", max(n, 1) * 8)
        self._alloc("D_OUT_W", max(n, 1) * 4)
        self._alloc("D_OUT_SUM_W", 8)

    def run(self):
        check(lib().icp_weights(self.ctx.h, self.buf["D_IN"].ptr, self.buf["D_OUT_W"].ptr, self.buf["D_OUT_SUM_W"].ptr, self.n))

    def read(self):
        return self.buf["D_OUT_W"].read(np.float32, self.n), float(self.buf["D_OUT_SUM_W"].read(np.float64, 1)[0])


class ICPMean(_Stage):
    """algorithms.hpp:624-837; config 'REGULAR' | 'WEIGHTED'"""

    def __init__(self, ctx, config="REGULAR"):
        super().__init__(ctx)
        self.config = config

    def init(self, n):
        self.n = n
        self._alloc("D_IN_F", max(n, 1) * 32)
        self._alloc("D_IN_M", max(n, 1) * 32)
        if self.config == "WEIGHTED":
            self._alloc("D_IN_W", max(n, 1) * 4)
            self._alloc("D_IN_SUM_W", 8)
        self._alloc("D_OUT", 32)

    def run(self):
        b = self.buf
        if self.config == "WEIGHTED":
            check(lib().icp_mean_weighted(self.ctx.h, b["D_IN_F"].ptr, b["D_IN_M"].ptr, b["D_IN_W"].ptr, b["D_IN_SUM_W"].ptr, b["D_OUT"].ptr, self.n))
        else:
            check(lib().icp_mean(self.ctx.h, b["D_IN_F"].ptr, b["D_IN_M"].ptr, b["D_OUT"].ptr, self.n))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, 8)


class ICPDevs(_Stage):
    """algorithms.hpp:867-939"""

    def init(self, n):
        self.n = n
        for name, nb in (("D_IN_F", n * 32), ("D_IN_M", n * 32), ("D_IN_MEAN", 32), ("D_OUT_DEV_F", n * 16), ("D_OUT_DEV_M", n * 16)):
            self._alloc(name, max(nb, 16))

    def run(self):
        b = self.buf
        check(lib().icp_devs(self.ctx.h, b["D_IN_F"].ptr, b["D_IN_M"].ptr, b["D_IN_MEAN"].ptr, b["D_OUT_DEV_F"].ptr, b["D_OUT_DEV_M"].ptr, self.n))

    def read(self):
        return self.buf["D_OUT_DEV_F"].read(np.float32, (self.n, 4)), self.buf["D_OUT_DEV_M"].read(np.float32, (self.n, 4))


class ICPS(_Stage):
    """algorithms.hpp:990-1185; config 'REGULAR' | 'WEIGHTED'"""

    def __init__(self, ctx, config="REGULAR"):
        super().__init__(ctx)
        self.config = config

    def init(self, m, c):
        self.m, self.c = m, c
        self._alloc("D_IN_DEV_M", max(m, 1) * 16)
        self._alloc("D_IN_DEV_F", max(m, 1) * 16)
        if self.config == "WEIGHTED":
            self._alloc("D_IN_W", max(m, 1) * 4)
        self._alloc("D_OUT", 64)

    def run(self):
        b = self.buf
        w = b["D_IN_W"].ptr if self.config == "WEIGHTED" else None
        check(lib().icp_sij(self.ctx.h, b["D_IN_DEV_M"].ptr, b["D_IN_DEV_F"].ptr, w, b["D_OUT"].ptr, self.m, self.c))

    def read(self):
        return self.buf["D_OUT"].read(np.float32, 11)


class ICPPowerMethod(_Stage):
    """algorithms.hpp:1451-1537"""

    def init(self):
        self._alloc("D_IN_S", 64)
        self._alloc("D_IN_MEAN", 32)
        self._alloc("D_OUT_T_K", 32)

    def run(self):
        check(lib().icp_power_method(self.ctx.h, self.buf["D_IN_S"].ptr, self.buf["D_IN_MEAN"].ptr, self.buf["D_OUT_T_K"].ptr))

    def read(self):
        return self.buf["D_OUT_T_K"].read(np.float32, 8)


class ICPSVD(_Stage):
    """The host Eigen SVD block of ICPStep<EIGEN,*>::run (algorithms.cpp:3877-3896), on the device."""

    def init(self):
        self._alloc("D_IN_S", 64)
        self._alloc("D_IN_MEAN", 32)
        self._alloc("D_OUT_T_K", 32)
        self._alloc("D_OUT_R_K", 48)

    def run(self):
        b = self.buf
        check(lib().icp_svd_solve(self.ctx.h, b["D_IN_S"].ptr, b["D_IN_MEAN"].ptr, b["D_OUT_T_K"].ptr, b["D_OUT_R_K"].ptr))

    def read(self):
        return self.buf["D_OUT_T_K"].read(np.float32, 8), self.buf["D_OUT_R_K"].read(np.float32, 9).reshape(3, 3)


class Reduce(_Stage):
    """algorithms.hpp:83-185; config 'MIN' (float) | 'MAX' (uint) | 'SUM' (float)"""

    def __init__(self, ctx, config):
        super().__init__(ctx)
        self.config = config

    def init(self, cols, rows):
        self.cols, self.rows = cols, rows
        self._alloc("D_IN", max(cols * rows, 1) * 4)
        self._alloc("D_OUT", max(rows, 1) * 4)

    def run(self):
        fn = {"MIN": lib().icp_reduce_min_f, "MAX": lib().icp_reduce_max_ui, "SUM": lib().icp_reduce_sum_f}[self.config]
        check(fn(self.ctx.h, self.buf["D_IN"].ptr, self.cols, self.rows, self.buf["D_OUT"].ptr))

    def read(self):
        return self.buf["D_OUT"].read(np.uint32 if self.config == "MAX" else np.float32, self.rows)


class Scan(_Stage):
    """algorithms.hpp:207-289; config 'INCLUSIVE' | 'EXCLUSIVE'"""

    def __init__(self, ctx, config):
        super().__init__(ctx)
        self.config = config

    def init(self, cols, rows):
        self.cols, self.rows = cols, rows
        self._alloc("D_IN", max(cols * rows, 1) * 4)
        self._alloc("D_OUT", max(cols * rows, 1) * 4)

    def run(self):
        check(lib().icp_scan_i(self.ctx.h, self.buf["D_IN"].ptr, self.cols, self.rows, 1 if self.config == "INCLUSIVE" else 0, self.buf["D_OUT"].ptr))

    def read(self):
        return self.buf["D_OUT"].read(np.int32, (self.rows, self.cols))


class ICPStep:
    """ICPStep<CR,CW> (algorithms.hpp:1613-2401): one registration engine; pose accumulated on the device."""

    def __init__(self, ctx, rot=ROT_POWER_METHOD, weighting=W_WEIGHTED):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib().icp_step_create(ctx.h, rot, weighting, C.byref(h)))
        self.h = h
        self.m = self.nr = 0

    def close(self):
        if self.h:
            lib().icp_step_destroy(self.h)
            self.h = None

    def bind(self, mem, devbuf):
        check(lib().icp_step_bind(self.h, mem, devbuf.ptr if isinstance(devbuf, DeviceBuffer) else devbuf))

    def init(self, m, nr, a=1e2, c=1e-6, lm_w=0, lm_h=0):
        check(lib().icp_step_init(self.h, m, nr, a, c, lm_w, lm_h))
        self.m, self.nr = m, nr

    def write(self, mem, arr, block=True):
        arr = np.ascontiguousarray(arr, np.float32)
        check(lib().icp_step_write(self.h, mem, arr.ctypes.data, 1 if block else 0))

    def buffer(self, mem):
        return lib().icp_step_buffer(self.h, mem)

    def reset(self):
        check(lib().icp_step_reset(self.h))

    def setAlpha(self, a):
        check(lib().icp_step_set_alpha(self.h, a))

    def setScaling(self, c):
        check(lib().icp_step_set_scaling(self.h, c))

    def set_metric(self, fg, fp):
        check(lib().icp_step_set_metric(self.h, fg, fp))

    def set_mode(self, mode):
        check(lib().icp_step_set_mode(self.h, mode))

    def set_count_evals(self, on):
        check(lib().icp_step_set_count_evals(self.h, 1 if on else 0))

    def buildRBC(self):
        check(lib().icp_step_build_rbc(self.h))

    def run(self, n_iters=1, variant=None):
        if variant is None:
            check(lib().icp_step_run(self.h, n_iters))
        else:
            check(lib().icp_step_run_variant(self.h, n_iters, variant))

    def run_timed(self):
        out = np.zeros(7, np.float32)
        check(lib().icp_step_run_timed(self.h, out.ctypes.data))
        return dict(zip(("transform", "rbc_search", "weights", "means", "devs", "S", "solve"), out.tolist()))

    def state(self):
        st = capi.IcpState()
        check(lib().icp_step_get_state(self.h, C.byref(st)))
        g = lambda a: np.array(list(a), np.float32)
        return dict(Rk=g(st.Rk).reshape(3, 3), qk=g(st.qk), tk=g(st.tk), sk=np.float32(st.sk), R=g(st.R).reshape(3, 3),
                    q=g(st.q), t=g(st.t), s=np.float32(st.s), k=int(st.k), done=int(st.done))

    def pose_matrix(self):
        out = np.zeros(16, np.float32)
        check(lib().icp_step_get_pose_matrix(self.h, out.ctypes.data))
        return out.reshape(4, 4)

    def eval_counts(self):
        a, b = C.c_uint64(), C.c_uint64()
        check(lib().icp_step_eval_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stage1_executed(self):
        a = C.c_uint64()
        check(lib().icp_step_stage1_executed(self.h, C.byref(a)))
        return a.value

    def stage2_executed(self):
        a = C.c_uint64()
        check(lib().icp_step_stage2_executed(self.h, C.byref(a)))
        return a.value

    def debug(self, name, dtype, shape):
        p = lib().icp_step_debug_ptr(self.h, name.encode())
        if not p:
            raise KeyError(name)
        return capi.read_ptr(self.ctx, p, dtype, shape)


class ICP(ICPStep):
    """ICP<CR,CW> (algorithms.hpp:2433-2496): the iterative driver with the convergence test on the device."""

    def init(self, m, nr, a=1e2, c=1e-6, max_iterations=40, angle_threshold=0.001, translation_threshold=0.01, lm_w=0, lm_h=0):
        self.max_iterations, self.angle_threshold, self.translation_threshold = max_iterations, angle_threshold, translation_threshold
        self.k = 0
        super().init(m, nr, a, c, lm_w, lm_h)

    def buildRBC(self):
        super().buildRBC()
        self.k = 0

    def run(self, n_iters=None, variant=None):
        """No argument: ICP::run() (blocking, convergence-driven).  With n_iters: fixed step count (profiling driver)."""
        if n_iters is not None:
            return super().run(n_iters, variant)
        k = C.c_uint32()
        check(lib().icp_run(self.h, self.max_iterations, self.angle_threshold, self.translation_threshold, C.byref(k)))
        self.k = k.value
        return self.k


class ICPBatch:
    """Independent frame pairs registered in lock step on one GPU (throughput mode)."""

    def __init__(self, ctx, n_pairs, m, nr, a=2e2, c=1e-6, rot=ROT_POWER_METHOD, weighting=W_WEIGHTED, lm_w=0, lm_h=0):
        self.ctx, self.n_pairs, self.m, self.nr = ctx, n_pairs, m, nr
        h = C.c_void_p()
        check(lib().icp_batch_create(ctx.h, rot, weighting, n_pairs, m, nr, a, c, lm_w, lm_h, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().icp_batch_destroy(self.h)
            self.h = None

    def synthesize(self, base_devbuf, seed):
        check(lib().icp_batch_synthesize(self.h, base_devbuf.ptr, seed))

    def upload(self, first, F, M, block=True):
        F = np.ascontiguousarray(F, np.float32)
        M = np.ascontiguousarray(M, np.float32)
        count = F.shape[0] if F.ndim == 3 else 1
        check(lib().icp_batch_upload(self.h, first, count, F.ctypes.data, M.ctypes.data, 1 if block else 0))

    def upload_ptr(self, first, count, hF, hM, block=False):
        check(lib().icp_batch_upload(self.h, first, count, hF, hM, 1 if block else 0))

    def register(self, n_iters):
        check(lib().icp_batch_register(self.h, n_iters))

    def set_slices(self, n_slices):
        """register() runs the batch as n_slices concurrent slices (same results)."""
        check(lib().icp_batch_set_slices(self.h, n_slices))

    def slices(self):
        return lib().icp_batch_slices(self.h)

    def cmode(self):
        """Kernel-C flavour: 0 = k_search<L>, 1 = k_search_grouped, 2 = k_colscan_sort + k_search_sorted."""
        return lib().icp_batch_cmode(self.h)

    def register_host(self, hF, hM, n_iters, n_slices=0):
        """Frames in host memory -> poses: sliced upload overlapped with the registration of the previous slice.
        hF / hM: numpy arrays [n_pairs, m, 8] (float32, contiguous) or raw host pointers (ints)."""
        if isinstance(hF, np.ndarray):
            assert hF.dtype == np.float32 and hF.flags.c_contiguous and hF.size == self.n_pairs * self.m * 8
            assert hM.dtype == np.float32 and hM.flags.c_contiguous and hM.size == self.n_pairs * self.m * 8
            hF, hM = hF.ctypes.data, hM.ctypes.data
        T8 = np.zeros((self.n_pairs, 8), np.float32)
        check(lib().icp_batch_register_host(self.h, hF, hM, n_iters, n_slices, T8.ctypes.data))
        return T8

    def register_host_async(self, hF, hM, n_iters, n_slices=0):
        """Enqueue uploads + registration + d2h of the poses on the batch's own streams; collect() waits for them."""
        if isinstance(hF, np.ndarray):
            assert hF.dtype == np.float32 and hF.flags.c_contiguous and hF.size == self.n_pairs * self.m * 8
            assert hM.dtype == np.float32 and hM.flags.c_contiguous and hM.size == self.n_pairs * self.m * 8
            self._keep = (hF, hM)                 # the host buffers must outlive the asynchronous copies
            hF, hM = hF.ctypes.data, hM.ctypes.data
        check(lib().icp_batch_register_host_async(self.h, hF, hM, n_iters, n_slices))

    def collect(self):
        T8 = np.zeros((self.n_pairs, 8), np.float32)
        check(lib().icp_batch_collect(self.h, T8.ctypes.data))
        self._keep = None
        return T8

    def read_poses(self, want_T16=False):
        T8 = np.zeros((self.n_pairs, 8), np.float32)
        T16 = np.zeros((self.n_pairs, 16), np.float32) if want_T16 else None
        check(lib().icp_batch_read_poses(self.h, T8.ctypes.data, T16.ctypes.data if want_T16 else None))
        return (T8, T16.reshape(-1, 4, 4)) if want_T16 else T8

    def time_kernel(self, which, n_launches=10):
        ms = C.c_float()
        check(lib().icp_batch_time_kernel(self.h, which, n_launches, C.byref(ms)))
        return ms.value

    def config(self):
        qb, nba, s, cl, l = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_int(), C.c_int()
        check(lib().icp_batch_config(self.h, C.byref(qb), C.byref(nba), C.byref(s), C.byref(cl), C.byref(l)))
        return dict(QB=qb.value, nbA=nba.value, S=s.value, CL=cl.value, L=l.value)

    def debug(self, name, dtype, shape, pair=0):
        p = lib().icp_batch_debug_ptr(self.h, f"{name}@{pair}".encode())
        if not p:
            raise KeyError(name)
        return capi.read_ptr(self.ctx, p, dtype, shape)


class ICPMulti:
    """Independent frame pairs registered on several GPUs of one box from ONE process (icp_multi_*): device d owns a
    contiguous block of pairs end to end, one host thread per device inside the call, no collective."""

    def __init__(self, n_pairs, m, nr, a=2e2, c=1e-6, rot=ROT_POWER_METHOD, weighting=W_WEIGHTED, n_devices=0, devices=None, lm_w=0, lm_h=0):
        self.n_pairs, self.m, self.nr = n_pairs, m, nr
        h = C.c_void_p()
        dev = None
        if devices is not None:
            dev = (C.c_int * len(devices))(*devices)
            n_devices = len(devices)
        check(lib().icp_multi_create(n_devices, dev, rot, weighting, n_pairs, m, nr, a, c, lm_w, lm_h, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().icp_multi_destroy(self.h)
            self.h = None

    def devices(self):
        return lib().icp_multi_devices(self.h)

    def pair_range(self, index):
        d, f, n = C.c_int(), C.c_uint32(), C.c_uint32()
        check(lib().icp_multi_pair_range(self.h, index, C.byref(d), C.byref(f), C.byref(n)))
        return d.value, f.value, n.value

    def register_host(self, hF, hM, n_iters):
        if isinstance(hF, np.ndarray):
            assert hF.dtype == np.float32 and hF.flags.c_contiguous and hF.size == self.n_pairs * self.m * 8
            assert hM.dtype == np.float32 and hM.flags.c_contiguous and hM.size == self.n_pairs * self.m * 8
            hF, hM = hF.ctypes.data, hM.ctypes.data
        T8 = np.zeros((self.n_pairs, 8), np.float32)
        check(lib().icp_multi_register_host(self.h, hF, hM, n_iters, T8.ctypes.data))
        return T8
