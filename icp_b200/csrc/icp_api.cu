// icp_api.cu -- C ABI glue: context, memory, timers, and the per-stage entry points of include/icp_b200.h.
#include "icp_stages.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void icp_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *icp_last_error(void) { return g_err; }
extern "C" const char *icp_version(void) { return "icp_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int icp_ctx_create(int device, void *cuda_stream, icp_ctx **out)
{
    if (!out) { icp_set_error("icp_ctx_create: out is NULL"); return ICP_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
    {
        icp_set_error("icp_ctx_create: no CUDA device available (%s); libicp_b200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return ICP_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { icp_set_error("icp_ctx_create: device %d out of range [0,%d)", device, ndev); return ICP_ERR_ARG; }
    IcpDeviceGuard guard__(device);
    ICP_CUDA(guard__.err);
    cudaDeviceProp prop;
    ICP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
    {
        icp_set_error("icp_ctx_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
        return ICP_ERR_CUDA;
    }
    icp_ctx *c = new icp_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major; c->cc_minor = prop.minor;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
    c->clock_khz = khz;
    // anything that fails from here on releases what was created (icp_ctx_destroy copes with a half-built context)
    cudaError_t ce = cudaSuccess;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else { ce = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking); c->own_stream = (ce == cudaSuccess); }
    if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&c->ev1);
    if (ce != cudaSuccess)
    {
        icp_set_error("CUDA error %s in icp_ctx_create: %s", cudaGetErrorName(ce), cudaGetErrorString(ce));
        icp_ctx_destroy(c);
        return ICP_ERR_CUDA;
    }
    *out = c;
    return ICP_OK;
}

extern "C" void icp_ctx_destroy(icp_ctx *ctx)
{
    if (!ctx) return;
    IcpDeviceGuard guard__(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int icp_ctx_sync(icp_ctx *ctx)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ICP_OK;
}

extern "C" int icp_device_info(icp_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz, size_t *l2_bytes)
{ ICP_ENTER(ctx);
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (clock_khz) *clock_khz = ctx->clock_khz;
    if (l2_bytes) *l2_bytes = ctx->l2_bytes;
    return ICP_OK;
}

int icp_ctx_scratch(icp_ctx *ctx, size_t bytes, void **out)
{
    if (bytes > ctx->scratch_bytes)
    {
        ICP_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) ICP_CUDA(cudaFree(ctx->scratch));
        ctx->scratch = nullptr; ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 2 + 4096;
        ICP_CUDA(cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// memory
// ------------------------------------------------------------------------------------------------
extern "C" int icp_malloc(icp_ctx *ctx, size_t bytes, void **d_ptr)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    return ICP_OK;
}
extern "C" int icp_free(icp_ctx *ctx, void *d_ptr)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaFree(d_ptr));
    return ICP_OK;
}
extern "C" int icp_host_alloc(size_t bytes, void **h_ptr) { ICP_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 16)); return ICP_OK; }
extern "C" int icp_host_free(void *h_ptr) { ICP_CUDA(cudaFreeHost(h_ptr)); return ICP_OK; }

extern "C" int icp_memcpy_h2d(icp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, int block)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (block) ICP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ICP_OK;
}
extern "C" int icp_memcpy_d2h(icp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, int block)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (block) ICP_CUDA(cudaStreamSynchronize(ctx->stream));
    return ICP_OK;
}
extern "C" int icp_memcpy_d2d(icp_ctx *ctx, void *d_dst, const void *d_src, size_t bytes)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return ICP_OK;
}
extern "C" int icp_memset(icp_ctx *ctx, void *d_dst, int value, size_t bytes)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// timing
// ------------------------------------------------------------------------------------------------
extern "C" int icp_timer_start(icp_ctx *ctx) { ICP_ENTER(ctx); ICP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream)); return ICP_OK; }
extern "C" int icp_timer_stop(icp_ctx *ctx, float *ms)
{ ICP_ENTER(ctx);
    ICP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    ICP_CUDA(cudaEventSynchronize(ctx->ev1));
    ICP_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return ICP_OK;
}

__global__ void k_fill(float4 *p, size_t n, float v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}

extern "C" int icp_flush_l2(icp_ctx *ctx)
{ ICP_ENTER(ctx);
    if (!ctx->l2_flush)
    {
        size_t bytes = ctx->l2_bytes * 2;
        if (bytes < ((size_t)256 << 20)) bytes = (size_t)256 << 20;
        ICP_CUDA(cudaMalloc(&ctx->l2_flush, bytes));
        ctx->l2_flush_bytes = bytes;
    }
    k_fill<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>((float4 *)ctx->l2_flush, ctx->l2_flush_bytes / 16, 1.f);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------------
#define REQUIRE(cond, cls, msg) do { if (!(cond)) ICP_CONFIG_FAIL(cls, msg); } while (0)

extern "C" int icp_get_lms(icp_ctx *ctx, const float *d_cloud, float *d_lms)
{ ICP_ENTER(ctx);
    REQUIRE(d_cloud && d_lms, "ICPLMs", "null buffer");
    return launch_get_lms(ctx->stream, d_cloud, d_lms);
}

extern "C" int icp_rgbd_to_pc8d(icp_ctx *ctx, const uint16_t *d_depth, const uint8_t *d_rgb, uint32_t W, uint32_t H, float focal, float *d_cloud)
{ ICP_ENTER(ctx);
    REQUIRE(d_depth && d_rgb && d_cloud, "RGBDTo8D", "null buffer");
    REQUIRE(W != 0 && H != 0 && (uint64_t)W * H <= (1u << 28), "RGBDTo8D", "The frame cannot have zero (or more than 2^28) pixels");
    REQUIRE(focal != 0.f, "RGBDTo8D", "The focal length cannot be zero");
    return launch_rgbd_to_pc8d(ctx->stream, ctx->sm_count, d_depth, d_rgb, W, H, focal, d_cloud);
}

extern "C" int icp_get_reps(icp_ctx *ctx, const float *d_lms, uint32_t W, uint32_t H, uint32_t nr, float *d_reps)
{ ICP_ENTER(ctx);
    REQUIRE(d_lms && d_reps, "ICPReps", "null buffer");
    REQUIRE(nr != 0, "ICPReps", "The array of representatives cannot have zero points");
    REQUIRE(nr % 4 == 0, "ICPReps", "The number of representatives has to be a multiple of 4");   // algorithms.cpp:842
    return launch_get_reps(ctx->stream, d_lms, W, H, nr, d_reps);
}

extern "C" int icp_transform_quaternion(icp_ctx *ctx, const float *d_M, const float *d_T8, float *d_out, uint32_t m)
{ ICP_ENTER(ctx);
    REQUIRE(m != 0, "ICPTransform<ICPTransformConfig::QUATERNION>", "The array cannot have zero points");
    return launch_transform_q(ctx->stream, d_M, d_T8, d_out, m);
}

extern "C" int icp_transform_matrix(icp_ctx *ctx, const float *d_M, const float *d_T16, float *d_out, uint32_t m)
{ ICP_ENTER(ctx);
    REQUIRE(m != 0, "ICPTransform<ICPTransformConfig::MATRIX>", "The array cannot have zero points");
    return launch_transform_m(ctx->stream, d_M, d_T16, d_out, m);
}

extern "C" int icp_rbc_construct(icp_ctx *ctx, const float *d_X, uint32_t n, const float *d_R, uint32_t nr, float alpha,
                                 uint32_t *d_rep_id, uint32_t *d_N, uint32_t *d_O, uint32_t *d_perm, float *d_Xp)
{ ICP_ENTER(ctx);
    REQUIRE(n != 0, "RBCConstruct", "The array X cannot have zero points");
    REQUIRE(nr != 0, "RBCConstruct", "The array R cannot have zero points");
    REQUIRE(alpha != 0.f, "RBCConstruct", "The alpha parameter cannot be equal to zero");
    REQUIRE(d_N && d_O && d_Xp, "RBCConstruct", "N, O and X_p outputs are mandatory");
    float fg, fp;
    icp_metric_weights(alpha, &fg, &fp);
    void *scr;
    const size_t sort_bytes = SortScratch::bytes(n, nr);
    ICP_CHECK(icp_ctx_scratch(ctx, sort_bytes + (size_t)n * 4 + 256, &scr));
    SortScratch sc; sc.carve(scr, n, nr);
    uint32_t *rep = d_rep_id ? d_rep_id : (uint32_t *)((char *)scr + ((sort_bytes + 255) & ~(size_t)255));
    ICP_CHECK(launch_nearest_rep(ctx->stream, ctx->sm_count, d_X, n, d_R, nr, fg, fp, nullptr, nullptr, rep));
    return launch_counting_sort(ctx->stream, rep, n, nr, sc, d_N, d_O, d_perm, d_X, d_Xp);
}

extern "C" int icp_rbc_search(icp_ctx *ctx, const float *d_Q, uint32_t m, const float *d_R, uint32_t nr, float alpha,
                              const float *d_Xp, const uint32_t *d_O, const uint32_t *d_N,
                              float *d_Qp, float *d_NN, icp_dist_id *d_NN_ID,
                              uint32_t *d_q_rep, uint32_t *d_qperm, uint32_t *d_Nq, uint32_t *d_Oq)
{ ICP_ENTER(ctx);
    REQUIRE(m != 0, "RBCSearch", "The array Q cannot have zero points");
    REQUIRE(nr != 0, "RBCSearch", "The array R cannot have zero points");
    REQUIRE(alpha != 0.f, "RBCSearch", "The alpha parameter cannot be equal to zero");
    REQUIRE(d_Qp && d_NN && d_NN_ID, "RBCSearch", "Q_p, NN and NN_ID outputs are mandatory");
    float fg, fp;
    icp_metric_weights(alpha, &fg, &fp);
    void *scr;
    const size_t sort_bytes = (SortScratch::bytes(m, nr) + 255) & ~(size_t)255;
    ICP_CHECK(icp_ctx_scratch(ctx, sort_bytes + (size_t)m * 8 + (size_t)nr * 8 + 1024, &scr));
    SortScratch sc; sc.carve(scr, m, nr);
    char *p = (char *)scr + sort_bytes;
    uint32_t *rep = d_q_rep ? d_q_rep : (uint32_t *)p; p += (size_t)m * 4;
    uint32_t *perm = d_qperm ? d_qperm : (uint32_t *)p; p += (size_t)m * 4;
    uint32_t *Nq = d_Nq ? d_Nq : (uint32_t *)p; p += (size_t)nr * 4;
    uint32_t *Oq = d_Oq ? d_Oq : (uint32_t *)p;
    ICP_CHECK(launch_nearest_rep(ctx->stream, ctx->sm_count, d_Q, m, d_R, nr, fg, fp, nullptr, nullptr, rep));
    ICP_CHECK(launch_counting_sort(ctx->stream, rep, m, nr, sc, Nq, Oq, perm, nullptr, nullptr));
    // database size = O[nr-1] + N[nr-1]; unknown on the host without a sync: pass an upper bound of 2^32-1 and
    // rely on lists being non-empty (R subset of X).  The kernel clamps only in the empty-list corner case.
    return launch_rbc_stage2(ctx->stream, d_Q, rep, perm, m, d_Xp, 0xFFFFFFFFu, d_O, d_N, fg, fp, d_Qp, d_NN, d_NN_ID, nullptr);
}

extern "C" int icp_weights(icp_ctx *ctx, const icp_dist_id *d_in, float *d_W, double *d_sum_w, uint32_t n)
{ ICP_ENTER(ctx);
    REQUIRE(n != 0, "ICPWeights", "The array cannot have zero elements");
    REQUIRE(n % 2 == 0, "ICPWeights", "The number of elements in the array must be a multiple of 2");   // algorithms.cpp:1049
    void *scr;
    const size_t e = reduce_scratch_elems(n);
    ICP_CHECK(icp_ctx_scratch(ctx, (e + 8) * 4 + 2 * e * 8 + 256, &scr));
    double *sd = (double *)scr;
    float *sf = (float *)((char *)scr + 2 * e * 8);
    return launch_weights(ctx->stream, d_in, d_W, d_sum_w, n, sf, sd);
}

extern "C" int icp_mean(icp_ctx *ctx, const float *d_F, const float *d_M, float *d_mean, uint32_t n)
{ ICP_ENTER(ctx);
    REQUIRE(n != 0, "ICPMean<ICPMeanConfig::REGULAR>", "The array cannot have zero points");
    REQUIRE(n % 2 == 0, "ICPMean<ICPMeanConfig::REGULAR>", "The number of points in the array must be a multiple of 2");
    void *scr;
    ICP_CHECK(icp_ctx_scratch(ctx, 16 * reduce_scratch_elems(n) * 4 + 256, &scr));
    return launch_mean(ctx->stream, d_F, d_M, nullptr, nullptr, d_mean, n, (float *)scr);
}

extern "C" int icp_mean_weighted(icp_ctx *ctx, const float *d_F, const float *d_M, const float *d_W, const double *d_sum_w,
                                 float *d_mean, uint32_t n)
{ ICP_ENTER(ctx);
    REQUIRE(n != 0, "ICPMean<ICPMeanConfig::WEIGHTED>", "The array cannot have zero points");
    REQUIRE(n % 2 == 0, "ICPMean<ICPMeanConfig::WEIGHTED>", "The number of points in the array must be a multiple of 2");
    REQUIRE(d_W && d_sum_w, "ICPMean<ICPMeanConfig::WEIGHTED>", "null weights buffer");
    void *scr;
    ICP_CHECK(icp_ctx_scratch(ctx, 16 * reduce_scratch_elems(n) * 4 + 256, &scr));
    return launch_mean(ctx->stream, d_F, d_M, d_W, d_sum_w, d_mean, n, (float *)scr);
}

extern "C" int icp_devs(icp_ctx *ctx, const float *d_F, const float *d_M, const float *d_mean, float *d_DF, float *d_DM, uint32_t n)
{ ICP_ENTER(ctx);
    REQUIRE(n != 0, "ICPDevs", "The array cannot have zero points");
    return launch_devs(ctx->stream, d_F, d_M, d_mean, d_DF, d_DM, n);
}

extern "C" int icp_sij(icp_ctx *ctx, const float *d_DM, const float *d_DF, const float *d_W, float *d_S11, uint32_t m, float c)
{ ICP_ENTER(ctx);
    REQUIRE(m != 0, "ICPS", "The array cannot have zero points");
    void *scr;
    ICP_CHECK(icp_ctx_scratch(ctx, (sij_partials_elems(m) + sij_scratch_elems(m)) * 4 + 256, &scr));
    float *part = (float *)scr;
    return launch_sij(ctx->stream, d_DM, d_DF, d_W, d_S11, m, c, part, part + sij_partials_elems(m));
}

extern "C" int icp_power_method(icp_ctx *ctx, const float *d_S11, const float *d_mean, float *d_Tk8)
{ ICP_ENTER(ctx);
    return launch_power_method(ctx->stream, d_S11, d_mean, d_Tk8);
}

extern "C" int icp_svd_solve(icp_ctx *ctx, const float *d_S11, const float *d_mean, float *d_Tk8, float *d_Rk9)
{ ICP_ENTER(ctx);
    return launch_svd_solve(ctx->stream, d_S11, d_mean, d_Tk8, d_Rk9);
}

extern "C" int icp_reduce_min_f(icp_ctx *ctx, const float *d_in, uint32_t cols, uint32_t rows, float *d_out)
{ ICP_ENTER(ctx);
    REQUIRE(cols != 0, "Reduce", "The array cannot have zero columns");
    REQUIRE(cols % 4 == 0, "Reduce", "The number of columns in the array must be a multiple of 4");   // algorithms.cpp:151
    return launch_reduce_min_f(ctx->stream, d_in, cols, rows, d_out);
}
extern "C" int icp_reduce_max_ui(icp_ctx *ctx, const uint32_t *d_in, uint32_t cols, uint32_t rows, uint32_t *d_out)
{ ICP_ENTER(ctx);
    REQUIRE(cols != 0, "Reduce", "The array cannot have zero columns");
    REQUIRE(cols % 4 == 0, "Reduce", "The number of columns in the array must be a multiple of 4");
    return launch_reduce_max_ui(ctx->stream, d_in, cols, rows, d_out);
}
extern "C" int icp_reduce_sum_f(icp_ctx *ctx, const float *d_in, uint32_t cols, uint32_t rows, float *d_out)
{ ICP_ENTER(ctx);
    REQUIRE(cols != 0, "Reduce", "The array cannot have zero columns");
    REQUIRE(cols % 4 == 0, "Reduce", "The number of columns in the array must be a multiple of 4");
    void *scr;
    ICP_CHECK(icp_ctx_scratch(ctx, reduce_sum_scratch_elems(cols, rows) * 4 + 256, &scr));
    return launch_reduce_sum_f(ctx->stream, d_in, cols, rows, d_out, (float *)scr);
}
extern "C" int icp_scan_i(icp_ctx *ctx, const int32_t *d_in, uint32_t cols, uint32_t rows, int inclusive, int32_t *d_out)
{ ICP_ENTER(ctx);
    REQUIRE(cols != 0, "Scan", "The array cannot have zero columns");
    return launch_scan_i(ctx->stream, d_in, cols, rows, inclusive, d_out);
}
