"""ctypes binding of libicp_b200.so (include/icp_b200.h).  No torch, no oracle, no CPU fallback: if the CUDA
extension is missing or no B200 is visible, every compute call raises.

Thin by design: numpy on the host side, raw device pointers on the device side.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ICP_B200_LIB: A/B-test another build of the same library (tools/tune*.py); never a fallback
LIB_PATH = os.environ.get("ICP_B200_LIB") or os.path.join(_HERE, "libicp_b200.so")

ICP_OK, ICP_ERR_CONFIG, ICP_ERR_CUDA, ICP_ERR_ARG = 0, 1, 2, 3
ROT_EIGEN, ROT_POWER_METHOD = 0, 1
W_REGULAR, W_WEIGHTED = 0, 1
MODE_STAGED, MODE_FUSED = 0, 1
MEM_D_IN_F, MEM_D_IN_M, MEM_D_IO_T = 3, 4, 5

vp = C.c_void_p
u32 = C.c_uint32
f32 = C.c_float


class ICPConfigError(ValueError):
    """The reference prints `Error[<Class>]: ...` and exits (algorithms.cpp:164-168)."""


class ICPCudaError(RuntimeError):
    """The reference lets cl::Error propagate."""


class IcpState(C.Structure):
    _fields_ = [("Rk", f32 * 9), ("qk", f32 * 4), ("tk", f32 * 3), ("sk", f32),
                ("R", f32 * 9), ("q", f32 * 4), ("t", f32 * 3), ("s", f32), ("k", u32), ("done", u32)]


# name -> (restype, argtypes); every symbol declared in include/icp_b200.h
SIGNATURES = {
    "icp_last_error": (C.c_char_p, []),
    "icp_version": (C.c_char_p, []),
    "icp_ctx_create": (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
    "icp_ctx_destroy": (None, [vp]),
    "icp_ctx_sync": (C.c_int, [vp]),
    "icp_device_info": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "icp_malloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "icp_free": (C.c_int, [vp, vp]),
    "icp_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "icp_host_free": (C.c_int, [vp]),
    "icp_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_int]),
    "icp_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t, C.c_int]),
    "icp_memcpy_d2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "icp_memset": (C.c_int, [vp, vp, C.c_int, C.c_size_t]),
    "icp_timer_start": (C.c_int, [vp]),
    "icp_timer_stop": (C.c_int, [vp, C.POINTER(f32)]),
    "icp_flush_l2": (C.c_int, [vp]),
    "icp_get_lms": (C.c_int, [vp, vp, vp]),
    "icp_rgbd_to_pc8d": (C.c_int, [vp, vp, vp, u32, u32, f32, vp]),
    "icp_get_reps": (C.c_int, [vp, vp, u32, u32, u32, vp]),
    "icp_transform_quaternion": (C.c_int, [vp, vp, vp, vp, u32]),
    "icp_transform_matrix": (C.c_int, [vp, vp, vp, vp, u32]),
    "icp_rbc_construct": (C.c_int, [vp, vp, u32, vp, u32, f32, vp, vp, vp, vp, vp]),
    "icp_rbc_search": (C.c_int, [vp, vp, u32, vp, u32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "icp_rbc_search_exact": (C.c_int, [vp, vp, u32, vp, u32, f32, vp, vp, vp, vp, vp, vp]),
    "icp_weights": (C.c_int, [vp, vp, vp, vp, u32]),
    "icp_mean": (C.c_int, [vp, vp, vp, vp, u32]),
    "icp_mean_weighted": (C.c_int, [vp, vp, vp, vp, vp, vp, u32]),
    "icp_devs": (C.c_int, [vp, vp, vp, vp, vp, vp, u32]),
    "icp_sij": (C.c_int, [vp, vp, vp, vp, vp, u32, f32]),
    "icp_power_method": (C.c_int, [vp, vp, vp, vp]),
    "icp_svd_solve": (C.c_int, [vp, vp, vp, vp, vp]),
    "icp_reduce_min_f": (C.c_int, [vp, vp, u32, u32, vp]),
    "icp_reduce_max_ui": (C.c_int, [vp, vp, u32, u32, vp]),
    "icp_reduce_sum_f": (C.c_int, [vp, vp, u32, u32, vp]),
    "icp_scan_i": (C.c_int, [vp, vp, u32, u32, C.c_int, vp]),
    "icp_step_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "icp_step_destroy": (None, [vp]),
    "icp_step_bind": (C.c_int, [vp, C.c_int, vp]),
    "icp_step_init": (C.c_int, [vp, u32, u32, f32, f32, u32, u32]),
    "icp_step_buffer": (vp, [vp, C.c_int]),
    "icp_step_write": (C.c_int, [vp, C.c_int, vp, C.c_int]),
    "icp_step_reset": (C.c_int, [vp]),
    "icp_step_set_alpha": (C.c_int, [vp, f32]),
    "icp_step_set_scaling": (C.c_int, [vp, f32]),
    "icp_step_set_metric": (C.c_int, [vp, f32, f32]),
    "icp_step_set_mode": (C.c_int, [vp, C.c_int]),
    "icp_step_build_rbc": (C.c_int, [vp]),
    "icp_step_run": (C.c_int, [vp, u32]),
    "icp_step_run_variant": (C.c_int, [vp, u32, C.c_int]),
    "icp_run": (C.c_int, [vp, u32, C.c_double, C.c_double, C.POINTER(u32)]),
    "icp_step_get_state": (C.c_int, [vp, C.POINTER(IcpState)]),
    "icp_step_get_pose_matrix": (C.c_int, [vp, vp]),
    "icp_step_debug_ptr": (vp, [vp, C.c_char_p]),
    "icp_step_run_timed": (C.c_int, [vp, vp]),
    "icp_step_set_count_evals": (C.c_int, [vp, C.c_int]),
    "icp_step_eval_counts": (C.c_int, [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "icp_step_stage1_executed": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "icp_step_stage2_executed": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
    "icp_batch_create": (C.c_int, [vp, C.c_int, C.c_int, u32, u32, u32, f32, f32, u32, u32, C.POINTER(vp)]),
    "icp_batch_destroy": (None, [vp]),
    "icp_batch_F": (vp, [vp]),
    "icp_batch_M": (vp, [vp]),
    "icp_batch_synthesize": (C.c_int, [vp, vp, C.c_uint64]),
    "icp_batch_upload": (C.c_int, [vp, u32, u32, vp, vp, C.c_int]),
    "icp_batch_register": (C.c_int, [vp, u32]),
    "icp_batch_set_slices": (C.c_int, [vp, u32]),
    "icp_batch_slices": (u32, [vp]),
    "icp_batch_cmode": (C.c_int, [vp]),
    "icp_batch_register_host": (C.c_int, [vp, vp, vp, u32, u32, vp]),
    "icp_batch_register_host_async": (C.c_int, [vp, vp, vp, u32, u32]),
    "icp_batch_collect": (C.c_int, [vp, vp]),
    "icp_batch_read_poses": (C.c_int, [vp, vp, vp]),
    "icp_batch_debug_ptr": (vp, [vp, C.c_char_p]),
    "icp_batch_time_kernel": (C.c_int, [vp, C.c_int, u32, C.POINTER(f32)]),
    "icp_batch_config": (C.c_int, [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "icp_multi_create": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, u32, u32, u32, f32, f32, u32, u32, C.POINTER(vp)]),
    "icp_multi_destroy": (None, [vp]),
    "icp_multi_devices": (C.c_int, [vp]),
    "icp_multi_pair_range": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(u32), C.POINTER(u32)]),
    "icp_multi_register_host": (C.c_int, [vp, vp, vp, u32, vp]),
    "icp_multi_register_host_once": (C.c_int, [C.c_int, C.c_int, C.c_int, u32, u32, u32, f32, f32, vp, vp, u32, vp]),
    "icp_measure_fp32_peak": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "icp_measure_fp32_rates": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "icp_measure_launch_floor": (C.c_int, [vp, C.POINTER(f32), C.POINTER(f32)]),
}

_LIB = None


def lib():
    """Load libicp_b200.so.  Raises if the extension was not built (no silent fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C icp_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)        # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc == ICP_OK:
        return
    msg = lib().icp_last_error().decode()
    if rc == ICP_ERR_CONFIG:
        raise ICPConfigError(msg)
    if rc == ICP_ERR_ARG:
        raise ValueError(msg)
    raise ICPCudaError(msg)


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, DeviceBuffer):
        return x.ptr
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x


class Context:
    """One device + one in-order stream (replaces clutils::CLEnv + CLEnvInfo<1>)."""

    def __init__(self, device=0, stream=None):
        h = vp()
        check(lib().icp_ctx_create(device, stream, C.byref(h)))
        self.h = h
        self.device = device
        sm, maj, mnr, khz, l2 = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        check(lib().icp_device_info(h, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(khz), C.byref(l2)))
        self.sm_count, self.cc, self.clock_khz, self.l2_bytes = sm.value, (maj.value, mnr.value), khz.value, l2.value

    def close(self):
        if self.h:
            lib().icp_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(lib().icp_ctx_sync(self.h))

    def timer_start(self):
        check(lib().icp_timer_start(self.h))

    def timer_stop(self):
        ms = f32()
        check(lib().icp_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        check(lib().icp_flush_l2(self.h))

    def alloc(self, nbytes):
        return DeviceBuffer(self, nbytes)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        b = DeviceBuffer(self, arr.nbytes)
        b.write(arr)
        return b


class DeviceBuffer:
    def __init__(self, ctx, nbytes, ptr=None):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        self.owned = ptr is None
        if ptr is None:
            p = vp()
            check(lib().icp_malloc(ctx.h, self.nbytes, C.byref(p)))
            ptr = p.value
        self.ptr = ptr

    def write(self, arr, block=True):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes, (arr.nbytes, self.nbytes)
        check(lib().icp_memcpy_h2d(self.ctx.h, self.ptr, arr.ctypes.data, arr.nbytes, 1 if block else 0))

    def read(self, dtype, shape):
        out = np.empty(shape, dtype)
        assert out.nbytes <= self.nbytes, (out.nbytes, self.nbytes)
        check(lib().icp_memcpy_d2h(self.ctx.h, out.ctypes.data, self.ptr, out.nbytes, 1))
        return out

    def free(self):
        if self.owned and self.ptr:
            lib().icp_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def read_ptr(ctx, ptr, dtype, shape):
    out = np.empty(shape, dtype)
    check(lib().icp_memcpy_d2h(ctx.h, out.ctypes.data, ptr, out.nbytes, 1))
    return out


class PinnedArray:
    """numpy view over pinned host memory (CL_MEM_ALLOC_HOST_PTR staging buffer of the reference)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = vp()
        check(lib().icp_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().icp_host_free(self.ptr)
            self.ptr = None
