// icp_bench.cu -- micro-benchmarks behind the roofline denominators reported by bench.py:
//   * FP32 CUDA-core issue rate for NON-fused mul/add (the distance path may not contract into FMA),
//     scalar (FMUL/FADD) and packed (FMUL2/FADD2, sm_100 f32x2);
//   * launch / graph-node latency floor.
#include "icp_common.cuh"

typedef unsigned long long u64;

template <int MODE>   // 0 scalar mul/add, 1 packed mul2/add2, 2 scalar FFMA, 3 packed FFMA2
__global__ void __launch_bounds__(256) k_fp32_rate(float *out, int iters, float x, float y)
{
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float b0 = a0 + .5f, b1 = a1 + .5f, b2 = a2 + .5f, b3 = a3 + .5f;
    if (MODE == 0)
    {
        for (int i = 0; i < iters; ++i)
        {
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                a0 = __fmul_rn(a0, x); a1 = __fmul_rn(a1, x); a2 = __fmul_rn(a2, x); a3 = __fmul_rn(a3, x);
                b0 = __fadd_rn(b0, y); b1 = __fadd_rn(b1, y); b2 = __fadd_rn(b2, y); b3 = __fadd_rn(b3, y);
            }
        }
    }
    else if (MODE == 2)
    {
        for (int i = 0; i < iters; ++i)
        {
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                a0 = __fmaf_rn(a0, x, y); a1 = __fmaf_rn(a1, x, y); a2 = __fmaf_rn(a2, x, y); a3 = __fmaf_rn(a3, x, y);
                b0 = __fmaf_rn(b0, x, y); b1 = __fmaf_rn(b1, x, y); b2 = __fmaf_rn(b2, x, y); b3 = __fmaf_rn(b3, x, y);
            }
        }
    }
    else
    {
        u64 A0, A1, A2, A3, B0, B1, B2, B3, X, Y;
        asm("mov.b64 %0, {%1,%2};" : "=l"(A0) : "f"(a0), "f"(a1));
        asm("mov.b64 %0, {%1,%2};" : "=l"(A1) : "f"(a2), "f"(a3));
        asm("mov.b64 %0, {%1,%2};" : "=l"(A2) : "f"(b0), "f"(b1));
        asm("mov.b64 %0, {%1,%2};" : "=l"(A3) : "f"(b2), "f"(b3));
        asm("mov.b64 %0, {%1,%2};" : "=l"(B0) : "f"(a1), "f"(a0));
        asm("mov.b64 %0, {%1,%2};" : "=l"(B1) : "f"(a3), "f"(a2));
        asm("mov.b64 %0, {%1,%2};" : "=l"(B2) : "f"(b1), "f"(b0));
        asm("mov.b64 %0, {%1,%2};" : "=l"(B3) : "f"(b3), "f"(b2));
        asm("mov.b64 %0, {%1,%2};" : "=l"(X) : "f"(x), "f"(x));
        asm("mov.b64 %0, {%1,%2};" : "=l"(Y) : "f"(y), "f"(y));
        for (int i = 0; i < iters; ++i)
        {
#pragma unroll
            for (int u = 0; u < 8; ++u)
            {
                if (MODE == 1)
                {
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A0) : "l"(X));
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A1) : "l"(X));
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A2) : "l"(X));
                    asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A3) : "l"(X));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(B0) : "l"(Y));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(B1) : "l"(Y));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(B2) : "l"(Y));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(B3) : "l"(Y));
                }
                else
                {
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A0) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A1) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A2) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A3) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(B0) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(B1) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(B2) : "l"(X), "l"(Y));
                    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(B3) : "l"(X), "l"(Y));
                }
            }
        }
        float t0, t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(A0)); a0 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(A1)); a1 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(A2)); a2 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(A3)); a3 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(B0)); b0 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(B1)); b1 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(B2)); b2 = t0 + t1;
        asm("mov.b64 {%0,%1}, %2;" : "=f"(t0), "=f"(t1) : "l"(B3)); b3 = t0 + t1;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((b0 + b1) + (b2 + b3));
}

template <int MODE>
static int time_rate(icp_ctx *ctx, float *buf, int grid, int iters, double *ops_per_s)
{
    float ms = 0.f, best = 1e30f;
    for (int rep = 0; rep < 4; ++rep)
    {
        ICP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
        k_fp32_rate<MODE><<<grid, 256, 0, ctx->stream>>>(buf, iters, 1.0000001f, 1e-7f);
        ICP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
        ICP_CUDA(cudaEventSynchronize(ctx->ev1));
        ICP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep && ms < best) best = ms;
    }
    // per thread per inner step: 8 instructions x 8 unroll; packed instructions carry 2 lanes
    const double lanes = (MODE == 1 || MODE == 3) ? 2.0 : 1.0;
    const double flop_per_instr = (MODE >= 2) ? 2.0 : 1.0;
    *ops_per_s = (double)grid * 256.0 * iters * 64.0 * lanes * flop_per_instr / (best * 1e-3);
    return ICP_OK;
}

// out[0] = scalar mul/add flop/s, out[1] = packed mul2/add2 flop/s, out[2] = scalar FFMA flop/s, out[3] = FFMA2 flop/s
extern "C" int icp_measure_fp32_rates(icp_ctx *ctx, double *out4)
{ ICP_ENTER(ctx);
    const int grid = ctx->sm_count * 8;
    float *buf = nullptr;
    ICP_CUDA(cudaMalloc((void **)&buf, (size_t)grid * 256 * sizeof(float)));
    const int iters = 4096;
    int rc = time_rate<0>(ctx, buf, grid, iters, &out4[0]);
    if (rc == ICP_OK) rc = time_rate<1>(ctx, buf, grid, iters, &out4[1]);
    if (rc == ICP_OK) rc = time_rate<2>(ctx, buf, grid, iters, &out4[2]);
    if (rc == ICP_OK) rc = time_rate<3>(ctx, buf, grid, iters, &out4[3]);
    cudaFree(buf);
    return rc;
}

extern "C" int icp_measure_fp32_peak(icp_ctx *ctx, double *flops_scalar, double *flops_packed)
{ ICP_ENTER(ctx);
    double r[4];
    ICP_CHECK(icp_measure_fp32_rates(ctx, r));
    if (flops_scalar) *flops_scalar = r[0];
    if (flops_packed) *flops_packed = r[1];
    return ICP_OK;
}

__global__ void k_empty(int *p) { if (p && threadIdx.x == 1024) *p = 0; }

extern "C" int icp_measure_launch_floor(icp_ctx *ctx, float *us_stream_launch, float *us_graph_node)
{ ICP_ENTER(ctx);
    cudaStream_t st = ctx->stream;
    const int n = 200;
    float ms = 0.f;
    for (int i = 0; i < 20; ++i) k_empty<<<1, 32, 0, st>>>(nullptr);
    ICP_CUDA(cudaStreamSynchronize(st));
    ICP_CUDA(cudaEventRecord(ctx->ev0, st));
    for (int i = 0; i < n; ++i) k_empty<<<1, 32, 0, st>>>(nullptr);
    ICP_CUDA(cudaEventRecord(ctx->ev1, st));
    ICP_CUDA(cudaEventSynchronize(ctx->ev1));
    ICP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (us_stream_launch) *us_stream_launch = ms * 1e3f / n;
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    ICP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < n; ++i) k_empty<<<1, 32, 0, st>>>(nullptr);
    ICP_CUDA(cudaStreamEndCapture(st, &g));
    ICP_CUDA(cudaGraphInstantiate(&ex, g, 0));
    ICP_CUDA(cudaGraphLaunch(ex, st));
    ICP_CUDA(cudaStreamSynchronize(st));
    ICP_CUDA(cudaEventRecord(ctx->ev0, st));
    ICP_CUDA(cudaGraphLaunch(ex, st));
    ICP_CUDA(cudaEventRecord(ctx->ev1, st));
    ICP_CUDA(cudaEventSynchronize(ctx->ev1));
    ICP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (us_graph_node) *us_graph_node = ms * 1e3f / n;
    cudaGraphExecDestroy(ex);
    cudaGraphDestroy(g);
    return ICP_OK;
}
