// icp_common.cuh -- shared host/device definitions of libicp_b200.so (sm_100a only).
//
// Floating-point contract (identical to oracle/icp_oracle.cpp): every f32 op is individually rounded,
// evaluation order as written.  The library is compiled with -fmad=false and the distance / reduction
// code additionally uses the __f*_rn intrinsics, which nvcc never contracts into FFMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/icp_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libicp_b200 is written for sm_100a (B200) only"
#endif

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void icp_set_error(const char *fmt, ...);

#define ICP_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            icp_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__,  \
                          cudaGetErrorString(e__));                                                 \
            return ICP_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define ICP_CHECK(expr)                                                                             \
    do { int r__ = (expr); if (r__ != ICP_OK) return r__; } while (0)

#define ICP_CONFIG_FAIL(cls, msg)                                                                   \
    do { icp_set_error("Error[%s]: %s", cls, msg); return ICP_ERR_CONFIG; } while (0)

#define ICP_LAUNCH_CHECK() ICP_CUDA(cudaGetLastError())

struct icp_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    int clock_khz = 0;
    size_t l2_bytes = 0;
    // scratch arena for the stage entry points (grown on demand, never during graph capture)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void *l2_flush = nullptr;
    size_t l2_flush_bytes = 0;
};

// Device guard (a process may hold contexts on several GPUs, and the host application's current device is its own
// business): every entry point that launches, copies, allocates or touches a stream / event / graph runs with the
// context's device current and restores the caller's device on exit.  device < 0 = no-op (null object: the entry
// point's own argument check reports it).
struct IcpDeviceGuard
{
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit IcpDeviceGuard(int dev)
    {
        if (dev < 0) return;
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~IcpDeviceGuard() { if (switched) cudaSetDevice(prev); }
    IcpDeviceGuard(const IcpDeviceGuard &) = delete;
    IcpDeviceGuard &operator=(const IcpDeviceGuard &) = delete;
};
#define ICP_ENTER(ctx)     IcpDeviceGuard guard__((ctx) ? (ctx)->device : -1); ICP_CUDA(guard__.err)
#define ICP_ENTER_OBJ(obj) IcpDeviceGuard guard__(((obj) && (obj)->ctx) ? (obj)->ctx->device : -1); ICP_CUDA(guard__.err)

int icp_ctx_scratch(icp_ctx *ctx, size_t bytes, void **out);

static inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// nr -> (nrx, nry) of ICPReps::init (algorithms.cpp:851-854)
static inline void icp_rep_grid(uint32_t nr, uint32_t *nrx, uint32_t *nry)
{
    int p = 0;
    while ((1u << (p + 1)) <= nr) ++p;
    *nrx = 1u << (p - p / 2);
    *nry = 1u << (p / 2);
}

// metric weights f_g(a), f_p(a) (decision B1), evaluated in f32 on the host exactly like the oracle
static inline void icp_metric_weights(float a, float *fg, float *fp)
{
    volatile float one_plus_a = 1.f + a;
    *fg = 1.f / one_plus_a;
    *fp = a / one_plus_a;
}

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

#define FULL_MASK 0xffffffffu

struct __align__(16) pt8 { float4 lo, hi; };

__device__ __forceinline__ pt8 ld_pt8(const float *base, uint32_t i)
{
    const float4 *p = reinterpret_cast<const float4 *>(base) + (size_t)i * 2;
    pt8 r; r.lo = __ldg(p); r.hi = __ldg(p + 1);
    return r;
}
__device__ __forceinline__ pt8 ld_pt8_cg(const float *base, uint32_t i)   // produced by an earlier kernel of the same graph
{
    const float4 *p = reinterpret_cast<const float4 *>(base) + (size_t)i * 2;
    pt8 r; r.lo = __ldcg(p); r.hi = __ldcg(p + 1);
    return r;
}
__device__ __forceinline__ void st_pt8(float *base, uint32_t i, const pt8 &v)
{
    float4 *p = reinterpret_cast<float4 *>(base) + (size_t)i * 2;
    p[0] = v.lo; p[1] = v.hi;
}

// 256-bit flavours (sm_100: LDG.E.ENL2.256 / STG.E.ENL2.256 -- one request and one L1 tag look-up per 32-byte point instead of
// two).  The address must be 32-byte aligned: used by the fused kernels on the engine's own buffers (icp_step_bind checks
// caller-provided ones); the stage entry points, which take arbitrary float4-aligned device pointers, keep the 128-bit pair.
__device__ __forceinline__ pt8 ld_pt8_v8(const float *base, uint32_t i)
{
    const float *p = base + (size_t)i * 8;
    pt8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p));
    return r;
}
__device__ __forceinline__ pt8 ld_pt8_cg_v8(const float *base, uint32_t i)
{
    const float *p = base + (size_t)i * 8;
    pt8 r;
    asm volatile("ld.global.cg.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_pt8_v8(float *base, uint32_t i, const pt8 &v)
{
    float *p = base + (size_t)i * 8;
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        :: "l"(p), "f"(v.lo.x), "f"(v.lo.y), "f"(v.lo.z), "f"(v.lo.w), "f"(v.hi.x), "f"(v.hi.y), "f"(v.hi.z), "f"(v.hi.w) : "memory");
}

// RBC metric (oracle dist8): fg*(((dx^2+dy^2)+dz^2)+dw^2) + fp*(((dr^2+dg^2)+db^2)+da^2)
__device__ __forceinline__ float dist8(const float4 &qlo, const float4 &qhi, const float4 &xlo, const float4 &xhi,
                                       float fg, float fp)
{
    float d0 = __fsub_rn(qlo.x, xlo.x), d1 = __fsub_rn(qlo.y, xlo.y), d2 = __fsub_rn(qlo.z, xlo.z), d3 = __fsub_rn(qlo.w, xlo.w);
    float d4 = __fsub_rn(qhi.x, xhi.x), d5 = __fsub_rn(qhi.y, xhi.y), d6 = __fsub_rn(qhi.z, xhi.z), d7 = __fsub_rn(qhi.w, xhi.w);
    float g = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)), __fmul_rn(d3, d3));
    float p = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d4, d4), __fmul_rn(d5, d5)), __fmul_rn(d6, d6)), __fmul_rn(d7, d7));
    return __fadd_rn(__fmul_rn(fg, g), __fmul_rn(fp, p));
}

// ICPTransform<QUATERNION> on one point (icp_kernels.cl:798; oracle orc_transform_q)
__device__ __forceinline__ float4 transform_q_xyz(const float4 &plo, const float4 &q, const float4 &t)
{
    const float px = plo.x, py = plo.y, pz = plo.z;
    const float q2x = __fmul_rn(2.f, q.x), q2y = __fmul_rn(2.f, q.y), q2z = __fmul_rn(2.f, q.z);
    // qcp = cross(q.xyz, p) + q.w * p
    float cx = __fsub_rn(__fmul_rn(q.y, pz), __fmul_rn(q.z, py));
    float cy = __fsub_rn(__fmul_rn(q.z, px), __fmul_rn(q.x, pz));
    float cz = __fsub_rn(__fmul_rn(q.x, py), __fmul_rn(q.y, px));
    cx = __fadd_rn(cx, __fmul_rn(q.w, px));
    cy = __fadd_rn(cy, __fmul_rn(q.w, py));
    cz = __fadd_rn(cz, __fmul_rn(q.w, pz));
    // c2 = cross(2 q.xyz, qcp)
    float ex = __fsub_rn(__fmul_rn(q2y, cz), __fmul_rn(q2z, cy));
    float ey = __fsub_rn(__fmul_rn(q2z, cx), __fmul_rn(q2x, cz));
    float ez = __fsub_rn(__fmul_rn(q2x, cy), __fmul_rn(q2y, cx));
    float4 o;
    o.x = __fadd_rn(__fmul_rn(t.w, __fadd_rn(px, ex)), t.x);
    o.y = __fadd_rn(__fmul_rn(t.w, __fadd_rn(py, ey)), t.y);
    o.z = __fadd_rn(__fmul_rn(t.w, __fadd_rn(pz, ez)), t.z);
    o.w = plo.w;
    return o;
}

// Reference work-group tree over a 128-slot block (icp_kernels.cl:169-175 at wgMultiple = 64):
// lane l of one warp holds slots l, l+32, l+64, l+96.  Result valid in lane 0.
__device__ __forceinline__ float warp_tree128(float e0, float e1, float e2, float e3)
{
    float a0 = __fadd_rn(e0, e2);            // d = 64: data[l] += data[l+64]   (l <  32)
    float a1 = __fadd_rn(e1, e3);            //                                   (l >= 32)
    float b = __fadd_rn(a0, a1);             // d = 32
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b = __fadd_rn(b, __shfl_down_sync(FULL_MASK, b, d));
    return b;
}
__device__ __forceinline__ double warp_tree128_d(double e0, double e1, double e2, double e3)
{
    double a0 = __dadd_rn(e0, e2);
    double a1 = __dadd_rn(e1, e3);
    double b = __dadd_rn(a0, a1);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) b = __dadd_rn(b, __shfl_down_sync(FULL_MASK, b, d));
    return b;
}

__device__ __forceinline__ float sum4f(const float4 &v)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(v.x, v.y), v.z), v.w);
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// device-resident registration state (public members of ICPStep, algorithms.hpp:2302-2320)
struct DevState
{
    float Rk[9], qk[4], tk[3], sk;
    float R[9], q[4], t[3], s;
    uint32_t k;
    uint32_t done;
};

#endif // __CUDACC__
