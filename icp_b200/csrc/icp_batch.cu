// icp_batch.cu -- batched registration of independent frame pairs on one GPU (throughput mode).
// All pairs advance in lock step through the fused kernels (grid.y = pair); one CUDA graph holds
// buildRBC + n iterations for the whole batch.  No collective, no host round trip; poses are read at the end.
#include "icp_fused.cuh"
#include <string.h>
#include <stdlib.h>

struct icp_batch
{
    icp_ctx *ctx = nullptr;
    uint32_t n_pairs = 0, m = 0, nr = 0, lm_w = 0, lm_h = 0;
    FusedCfg cfg;
    float *F = nullptr, *M = nullptr;        // [n_pairs][m][8]
    void *arena = nullptr;
    PairPtrs *table = nullptr;               // device [n_pairs]
    float *T = nullptr;                      // [n_pairs][8]
    DevState *state = nullptr;               // [n_pairs]
    LoopParams *loop = nullptr;              // [n_pairs]
    float *gt = nullptr;                     // [n_pairs][8] ground-truth poses of the synthesizer
    uint32_t *nn_id0 = nullptr;
    std::vector<PairPtrs> h_table;
    std::map<uint32_t, cudaGraphExec_t> graphs;
    std::map<uint64_t, cudaGraphExec_t> slice_graphs;       // (n_iters, first pair, count) -> graph over a slice of the table
    uint32_t n_slices = 1;                   // slices icp_batch_register runs concurrently (icp_batch_set_slices)
    cudaStream_t home_stream = nullptr;      // fork / join stream of the asynchronous host-buffer entry
    bool pending = false;                    // icp_batch_register_host_async enqueued, icp_batch_collect not yet called
    cudaStream_t copy_stream = nullptr;      // h2d uploads of slice i+1 run here while slice i registers
    std::vector<cudaStream_t> streams;       // compute streams of the slices
    std::vector<cudaEvent_t> ev_up, ev_done; // per slice: upload done; per stream: chain done
    cudaEvent_t ev_free = nullptr;           // compute stream has consumed the previous contents of F / M
    float *h_T = nullptr; icp_state *h_state = nullptr;
};

struct BatchLayout { size_t per_pair, total; };

static size_t carve_pair(Carver &cv, uint32_t m, uint32_t nr, uint32_t nbA, PairPtrs *P)
{
    PairPtrs q; memset(&q, 0, sizeof(q));
    q.reps = cv.take<float>((size_t)nr * 8);
    q.Xp = cv.take<float>((size_t)m * 8);
    q.N = cv.take<uint32_t>(nr); q.O = cv.take<uint32_t>(nr);
    q.rep_id = cv.take<uint32_t>(m); q.perm = cv.take<uint32_t>(m);
    q.q_rep = cv.take<uint32_t>(m);
    q.lrank = cv.take<uint16_t>((size_t)m + 8);
    q.H = cv.take<uint32_t>((size_t)nbA * nr + 32);
    q.Nq = cv.take<uint32_t>(nr); q.Oq = cv.take<uint32_t>(nr);
    q.wconst = cv.take<uint32_t>(16);
    q.nbr = cv.take<uint2>(fused_nbr_elems(nr));
    q.nbx = cv.take<uint32_t>((size_t)m * FUSED_NBX_K + 8);
    q.nn_o = cv.take<uint32_t>(m);
    q.nnd = cv.take<float>(m);
    q.qperm = cv.take<uint32_t>(m);
    q.QR = cv.take<uint2>(m);
    q.nn2 = cv.take<uint2>(m);
    q.Qs = cv.take<float>((size_t)m * 8);
    q.Rs = cv.take<uint4>(m);
    q.W = cv.take<float>(m);
    q.fxyz = cv.take<float>((size_t)3 * m); q.mxyz = cv.take<float>((size_t)3 * m);
    q.NNID = cv.take<icp_dist_id>(m);
    q.sum_w = cv.take<double>(2);
    q.mean = cv.take<float>(8); q.S = cv.take<float>(16); q.Tk = cv.take<float>(8); q.Rk = cv.take<float>(12);
    q.red = cv.take<float>(fused_red_elems(m));
    q.evals = cv.take<unsigned long long>(4);            // attached only when ICP_B200_BATCH_EVALS is set (diagnosis)
    q.prof = cv.take<unsigned long long>(64);            // attached only when ICP_B200_BATCH_PROF is set (phase clocks of kernel C')
    if (P) *P = q;
    return cv.off;
}

__global__ void k_batch_reset(DevState *state, float *T, LoopParams *loop, uint32_t n_pairs, int32_t iters)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    DevState *st = state + p;
    for (int i = 0; i < 9; ++i) { st->R[i] = (i % 4 == 0) ? 1.f : 0.f; st->Rk[i] = (i % 4 == 0) ? 1.f : 0.f; }
    for (int i = 0; i < 4; ++i) { st->q[i] = (i == 3) ? 1.f : 0.f; st->qk[i] = (i == 3) ? 1.f : 0.f; }
    for (int i = 0; i < 3; ++i) { st->t[i] = 0.f; st->tk[i] = 0.f; }
    st->s = 1.f; st->sk = 1.f; st->k = 0; st->done = 0;
    float *t = T + (size_t)p * 8;
    t[0] = 0.f; t[1] = 0.f; t[2] = 0.f; t[3] = 1.f; t[4] = 0.f; t[5] = 0.f; t[6] = 0.f; t[7] = 1.f;
    LoopParams lp; lp.angle_thr = 0.0; lp.trans_thr = 0.0; lp.check = 0; lp.max_iterations = 0; lp.iters_left = iters; lp.pad = 0;
    loop[p] = lp;
}

extern "C" void icp_batch_destroy(icp_batch *b);

// allocations of icp_batch_create; on failure the caller destroys the half-built object (no leak on error paths)
static int batch_create_impl(icp_batch *b, icp_ctx *ctx, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr,
                             float alpha, float c, uint32_t lm_w, uint32_t lm_h)
{
    b->ctx = ctx; b->n_pairs = n_pairs; b->m = m; b->nr = nr; b->lm_w = lm_w; b->lm_h = lm_h;
    // concurrent slices of >= 64 pairs (measured on B200, 256 pairs, us per pair-iteration: 1 slice 3.02, 2: 2.68, 4: 2.67, 8: 2.70)
    b->n_slices = n_pairs / 64u < 1u ? 1u : (n_pairs / 64u > 8u ? 8u : n_pairs / 64u);
    if (const char *e = getenv("ICP_B200_BATCH_SLICES")) if (atoi(e) > 0) b->n_slices = (uint32_t)atoi(e);
    fused_choose_cfg(&b->cfg, m, nr, ctx->sm_count, n_pairs);
    icp_metric_weights(alpha, &b->cfg.fg, &b->cfg.fp);
    b->cfg.c = c; b->cfg.weighted = w_cfg; b->cfg.power_method = (rot_cfg == ICP_ROT_POWER_METHOD);
    const size_t fm = (size_t)n_pairs * m * 8 * sizeof(float);
    ICP_CUDA(cudaMalloc((void **)&b->F, fm));
    ICP_CUDA(cudaMalloc((void **)&b->M, fm));
    Carver c0(nullptr);
    const size_t per_pair = (carve_pair(c0, m, nr, b->cfg.nbA, nullptr) + 255) & ~(size_t)255;
    Carver ch(nullptr);
    ch.take<PairPtrs>(n_pairs); ch.take<float>((size_t)n_pairs * 8); ch.take<DevState>(n_pairs); ch.take<LoopParams>(n_pairs);
    ch.take<float>((size_t)n_pairs * 8);
    const size_t head = (ch.off + 255) & ~(size_t)255;
    const size_t total = head + per_pair * n_pairs + 256;
    ICP_CUDA(cudaMalloc(&b->arena, total));
    ICP_CUDA(cudaMemsetAsync(b->arena, 0, total, ctx->stream));
    Carver cv(b->arena);
    b->table = cv.take<PairPtrs>(n_pairs);
    b->T = cv.take<float>((size_t)n_pairs * 8);
    b->state = cv.take<DevState>(n_pairs);
    b->loop = cv.take<LoopParams>(n_pairs);
    b->gt = cv.take<float>((size_t)n_pairs * 8);
    b->h_table.resize(n_pairs);
    for (uint32_t p = 0; p < n_pairs; ++p)
    {
        Carver cp((char *)b->arena + head + per_pair * p);
        PairPtrs P;
        carve_pair(cp, m, nr, b->cfg.nbA, &P);
        P.F = b->F + (size_t)p * m * 8; P.M = b->M + (size_t)p * m * 8;
        P.T = b->T + (size_t)p * 8; P.state = b->state + p; P.loop = b->loop + p;
        if (!getenv("ICP_B200_BATCH_EVALS")) P.evals = nullptr;
        if (!getenv("ICP_B200_BATCH_PROF")) P.prof = nullptr;
        b->h_table[p] = P;
    }
    ICP_CUDA(cudaMemcpyAsync(b->table, b->h_table.data(), sizeof(PairPtrs) * n_pairs, cudaMemcpyHostToDevice, ctx->stream));
    ICP_CUDA(cudaStreamSynchronize(ctx->stream));
    ICP_CUDA(cudaMallocHost((void **)&b->h_T, (size_t)n_pairs * 8 * sizeof(float)));
    ICP_CUDA(cudaMallocHost((void **)&b->h_state, (size_t)n_pairs * sizeof(icp_state)));
    return ICP_OK;
}

extern "C" int icp_batch_create(icp_ctx *ctx, int rot_cfg, int w_cfg, uint32_t n_pairs, uint32_t m, uint32_t nr,
                                float alpha, float c, uint32_t lm_w, uint32_t lm_h, icp_batch **out)
{ ICP_ENTER(ctx);
    if (!ctx || !out || n_pairs == 0) { icp_set_error("icp_batch_create: bad argument"); return ICP_ERR_ARG; }
    if (m == 0 || m > (1u << 20) || m % 2) ICP_CONFIG_FAIL("ICPBatch", "The sets of landmarks must have an even number of points in [2, 1048576]");
    if (nr == 0 || nr % 4 || nr > 4096) ICP_CONFIG_FAIL("ICPBatch", "The number of representatives has to be a multiple of 4 in [4, 4096]");
    if (alpha == 0.f) ICP_CONFIG_FAIL("ICPBatch", "The alpha parameter cannot be equal to zero");
    if (lm_w == 0 && lm_h == 0) { lm_w = 128; lm_h = 128; }
    if ((uint64_t)lm_w * lm_h != m) ICP_CONFIG_FAIL("ICPBatch", "The landmark grid (lm_w x lm_h) must hold exactly m points");
    uint32_t nrx, nry;
    icp_rep_grid(nr, &nrx, &nry);
    if (nrx * nry != nr || lm_w % nrx || lm_h % nry || lm_w / nrx < 2 || lm_h / nry < 2)
        ICP_CONFIG_FAIL("ICPReps", "The landmark grid is not divisible into the representative grid");
    icp_batch *b = new icp_batch();
    const int rc = batch_create_impl(b, ctx, rot_cfg, w_cfg, n_pairs, m, nr, alpha, c, lm_w, lm_h);
    if (rc != ICP_OK) { icp_batch_destroy(b); return rc; }      // releases whatever was allocated before the failure
    *out = b;
    return ICP_OK;
}

extern "C" void icp_batch_destroy(icp_batch *b)
{
    if (!b) return;
    IcpDeviceGuard guard__(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    for (auto &kv : b->graphs) cudaGraphExecDestroy(kv.second);
    for (auto &kv : b->slice_graphs) cudaGraphExecDestroy(kv.second);
    for (cudaEvent_t e : b->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : b->ev_done) cudaEventDestroy(e);
    for (cudaStream_t x : b->streams) { cudaStreamSynchronize(x); cudaStreamDestroy(x); }
    if (b->ev_free) cudaEventDestroy(b->ev_free);
    if (b->copy_stream) { cudaStreamSynchronize(b->copy_stream); cudaStreamDestroy(b->copy_stream); }
    if (b->home_stream) { cudaStreamSynchronize(b->home_stream); cudaStreamDestroy(b->home_stream); }
    if (b->F) cudaFree(b->F);
    if (b->M) cudaFree(b->M);
    if (b->arena) cudaFree(b->arena);
    if (b->h_T) cudaFreeHost(b->h_T);
    if (b->h_state) cudaFreeHost(b->h_state);
    delete b;
}

extern "C" float *icp_batch_F(icp_batch *b) { return b->F; }
extern "C" float *icp_batch_M(icp_batch *b) { return b->M; }

// ---- synthetic pairs generated on the device ----
__device__ __forceinline__ uint64_t splitmix(uint64_t &x)
{
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t &x) { return (float)((splitmix(x) >> 40) + 1) * (1.0f / 16777217.0f); }
__device__ __forceinline__ float gauss(uint64_t &x)
{
    const float u1 = u01(x), u2 = u01(x);
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

__global__ void k_synth(const float *__restrict__ base, float *__restrict__ F, float *__restrict__ M, float *__restrict__ gt,
                        uint32_t m, uint64_t seed)
{
    const uint32_t p = blockIdx.y;
    // pose of pair p: random axis, angle U[0.5,5] deg, |t| U[5,50] mm
    uint64_t sp = seed + 0x1000ull * p;
    float ax = gauss(sp), ay = gauss(sp), az = gauss(sp);
    const float an = rsqrtf(ax * ax + ay * ay + az * az);
    ax *= an; ay *= an; az *= an;
    const float ang = (0.5f + 4.5f * u01(sp)) * 0.017453292519943295f;
    float tx = gauss(sp), ty = gauss(sp), tz = gauss(sp);
    const float tn = (5.f + 45.f * u01(sp)) * rsqrtf(tx * tx + ty * ty + tz * tz);
    tx *= tn; ty *= tn; tz *= tn;
    const float sh = sinf(0.5f * ang), chf = cosf(0.5f * ang);
    const float qx = ax * sh, qy = ay * sh, qz = az * sh, qw = chf;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        float *g = gt + (size_t)p * 8;
        g[0] = qx; g[1] = qy; g[2] = qz; g[3] = qw; g[4] = tx; g[5] = ty; g[6] = tz; g[7] = 1.f;
    }
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    pt8 f = ld_pt8(base, i);
    st_pt8(F + (size_t)p * m * 8, i, f);
    uint64_t s = seed * 0x2545F4914F6CDD1Dull + ((uint64_t)p << 32) + i;
    pt8 mo = f;
    const bool valid = (f.lo.x != 0.f) || (f.lo.y != 0.f) || (f.lo.z != 0.f);
    if (valid)
    {
        // M = R^T (F - t) + noise : rotate by the conjugate quaternion
        const float px = f.lo.x - tx, py = f.lo.y - ty, pz = f.lo.z - tz;
        const float cx = -(qy * pz - qz * py) + qw * px, cy = -(qz * px - qx * pz) + qw * py, cz = -(qx * py - qy * px) + qw * pz;
        const float ex = -2.f * (qy * cz - qz * cy), ey = -2.f * (qz * cx - qx * cz), ez = -2.f * (qx * cy - qy * cx);
        mo.lo.x = px + ex + gauss(s); mo.lo.y = py + ey + gauss(s); mo.lo.z = pz + ez + gauss(s);
        if (u01(s) < 0.05f)
        {
            mo.lo.x = -2500.f + 5000.f * u01(s); mo.lo.y = -2000.f + 3500.f * u01(s); mo.lo.z = 500.f + 4000.f * u01(s);
            mo.hi.x = u01(s); mo.hi.y = u01(s); mo.hi.z = u01(s);
        }
    }
    mo.hi.x = fminf(1.f, fmaxf(0.f, mo.hi.x + 0.005f * gauss(s)));
    mo.hi.y = fminf(1.f, fmaxf(0.f, mo.hi.y + 0.005f * gauss(s)));
    mo.hi.z = fminf(1.f, fmaxf(0.f, mo.hi.z + 0.005f * gauss(s)));
    st_pt8(M + (size_t)p * m * 8, i, mo);
}

extern "C" int icp_batch_synthesize(icp_batch *b, const float *d_base, uint64_t seed)
{ ICP_ENTER_OBJ(b);
    k_synth<<<dim3(div_up(b->m, 256), b->n_pairs), 256, 0, b->ctx->stream>>>(d_base, b->F, b->M, b->gt, b->m, seed);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

extern "C" int icp_batch_upload(icp_batch *b, uint32_t first_pair, uint32_t count, const float *h_F, const float *h_M, int block)
{ ICP_ENTER_OBJ(b);
    if ((uint64_t)first_pair + count > b->n_pairs) { icp_set_error("icp_batch_upload: pair range out of bounds"); return ICP_ERR_ARG; }
    const size_t per = (size_t)b->m * 8 * sizeof(float);
    if (h_F) ICP_CUDA(cudaMemcpyAsync(b->F + (size_t)first_pair * b->m * 8, h_F, per * count, cudaMemcpyHostToDevice, b->ctx->stream));
    if (h_M) ICP_CUDA(cudaMemcpyAsync(b->M + (size_t)first_pair * b->m * 8, h_M, per * count, cudaMemcpyHostToDevice, b->ctx->stream));
    if (block) ICP_CUDA(cudaStreamSynchronize(b->ctx->stream));
    return ICP_OK;
}

static int batch_run_sliced(icp_batch *b, uint32_t n_iters, uint32_t n_slices, const float *h_F, const float *h_M, cudaStream_t home);

extern "C" int icp_batch_register(icp_batch *b, uint32_t n_iters)
{ ICP_ENTER_OBJ(b);
    if (n_iters == 0) return ICP_OK;
    cudaStream_t st = b->ctx->stream;
    // profiler aid: ncu does not list kernels that use the device-side graph API (kernel D) when they replay from a graph
    if (const char *e = getenv("ICP_B200_NO_GRAPH"))
        if (atoi(e) != 0)
        {
            k_batch_reset<<<div_up(b->n_pairs, 128), 128, 0, st>>>(b->state, b->T, b->loop, b->n_pairs, (int32_t)n_iters);
            ICP_CHECK(fused_launch_build(st, b->cfg, b->table, b->n_pairs, b->lm_w, b->lm_h));
            for (uint32_t i = 0; i < n_iters; ++i) ICP_CHECK(fused_launch_iteration(st, b->cfg, b->table, b->n_pairs, 0, 0));
            return ICP_OK;
        }
    if (b->n_slices > 1) return batch_run_sliced(b, n_iters, b->n_slices, nullptr, nullptr, nullptr);
    auto it = b->graphs.find(n_iters);
    cudaGraphExec_t ex = nullptr;
    if (it != b->graphs.end()) ex = it->second;
    else
    {
        cudaGraph_t g = nullptr;
        ICP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        int rc = ICP_OK;
        k_batch_reset<<<div_up(b->n_pairs, 128), 128, 0, st>>>(b->state, b->T, b->loop, b->n_pairs, (int32_t)n_iters);
        rc = fused_launch_build(st, b->cfg, b->table, b->n_pairs, b->lm_w, b->lm_h);
        for (uint32_t i = 0; i < n_iters && rc == ICP_OK; ++i) rc = fused_launch_iteration(st, b->cfg, b->table, b->n_pairs, 0, 0);
        cudaError_t e = cudaStreamEndCapture(st, &g);
        if (rc != ICP_OK) { if (g) cudaGraphDestroy(g); return rc; }
        ICP_CUDA(e);
        ICP_CUDA(cudaGraphInstantiate(&ex, g, 0));
        cudaGraphDestroy(g);
        if (b->graphs.size() >= 16)
        {
            ICP_CUDA(cudaStreamSynchronize(st));
            for (auto &kv : b->graphs) cudaGraphExecDestroy(kv.second);
            b->graphs.clear();
        }
        b->graphs[n_iters] = ex;
    }
    ICP_CUDA(cudaGraphLaunch(ex, st));
    return ICP_OK;
}

// ---- sliced execution -------------------------------------------------------------------------------------------
// The batch is cut into n_slices slices of consecutive pairs; every slice has its own graph (buildRBC + n_iters
// iterations over its part of the pair table) and the slices run on separate streams, so the latency-bound kernels of
// one slice (B, D: one wave of small CTAs) overlap the FP32-bound kernels (A, C) of the others, and -- with host
// buffers -- slice i+1 uploads on the copy stream while slice i computes.  Pairs are independent: results do not
// depend on the slicing.  Fork/join through events keeps the ordering of the context stream for the caller.
#define BATCH_MAX_STREAMS 8
static int batch_slice_graph(icp_batch *b, cudaStream_t cs, uint32_t n_iters, uint32_t first, uint32_t count, cudaGraphExec_t *out)
{
    const uint64_t key = ((uint64_t)n_iters << 48) | ((uint64_t)first << 24) | count;
    auto it = b->slice_graphs.find(key);
    if (it != b->slice_graphs.end()) { *out = it->second; return ICP_OK; }
    cudaGraph_t g = nullptr;
    cudaGraphExec_t ex = nullptr;
    ICP_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    k_batch_reset<<<div_up(count, 128), 128, 0, cs>>>(b->state + first, b->T + (size_t)first * 8, b->loop + first, count, (int32_t)n_iters);
    int rc = fused_launch_build(cs, b->cfg, b->table + first, count, b->lm_w, b->lm_h);
    for (uint32_t i = 0; i < n_iters && rc == ICP_OK; ++i) rc = fused_launch_iteration(cs, b->cfg, b->table + first, count, 0, 0);
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    if (rc != ICP_OK) { if (g) cudaGraphDestroy(g); return rc; }
    ICP_CUDA(e);
    ICP_CUDA(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
    if (b->slice_graphs.size() >= 64)   // bounded cache (distinct iteration counts x slicings); cached graphs may still be running
    {
        ICP_CUDA(cudaDeviceSynchronize());
        for (auto &kv : b->slice_graphs) cudaGraphExecDestroy(kv.second);
        b->slice_graphs.clear();
    }
    b->slice_graphs[key] = ex;
    *out = ex;
    return ICP_OK;
}

static int batch_run_sliced(icp_batch *b, uint32_t n_iters, uint32_t n_slices, const float *h_F, const float *h_M, cudaStream_t home)
{
    if (n_slices > b->n_pairs) n_slices = b->n_pairs;
    if (n_slices > 256 || n_iters >= (1u << 16) || b->n_pairs >= (1u << 24))
    { icp_set_error("icp_batch: at most 256 slices, 65535 iterations, 2^24 pairs"); return ICP_ERR_ARG; }
    cudaStream_t st = home ? home : b->ctx->stream;
    if (!b->copy_stream) ICP_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    if (!b->ev_free) ICP_CUDA(cudaEventCreateWithFlags(&b->ev_free, cudaEventDisableTiming));
    const uint32_t n_streams = n_slices < BATCH_MAX_STREAMS ? n_slices : BATCH_MAX_STREAMS;
    while (b->streams.size() < n_streams)
    {
        cudaStream_t x; ICP_CUDA(cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking));
        b->streams.push_back(x);
        cudaEvent_t e; ICP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->ev_done.push_back(e);
    }
    while (b->ev_up.size() < n_slices)
    {
        cudaEvent_t e; ICP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        b->ev_up.push_back(e);
    }
    // fork: earlier work on the context stream (uploads, previous registrations) comes first
    ICP_CUDA(cudaEventRecord(b->ev_free, st));
    if (h_F || h_M) ICP_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_free, 0));
    for (uint32_t x = 0; x < n_streams; ++x) ICP_CUDA(cudaStreamWaitEvent(b->streams[x], b->ev_free, 0));
    const size_t per = (size_t)b->m * 8;
    const uint32_t base = b->n_pairs / n_slices, extra = b->n_pairs % n_slices;
    uint32_t first = 0;
    for (uint32_t s = 0; s < n_slices; ++s)
    {
        const uint32_t count = base + (s < extra ? 1u : 0u);
        cudaStream_t cs = b->streams[s % n_streams];
        cudaGraphExec_t ex = nullptr;
        ICP_CHECK(batch_slice_graph(b, cs, n_iters, first, count, &ex));
        if (h_F || h_M)
        {
            if (h_F) ICP_CUDA(cudaMemcpyAsync(b->F + first * per, h_F + first * per, per * count * sizeof(float), cudaMemcpyHostToDevice, b->copy_stream));
            if (h_M) ICP_CUDA(cudaMemcpyAsync(b->M + first * per, h_M + first * per, per * count * sizeof(float), cudaMemcpyHostToDevice, b->copy_stream));
            ICP_CUDA(cudaEventRecord(b->ev_up[s], b->copy_stream));
            ICP_CUDA(cudaStreamWaitEvent(cs, b->ev_up[s], 0));
        }
        ICP_CUDA(cudaGraphLaunch(ex, cs));
        first += count;
    }
    // join
    for (uint32_t x = 0; x < n_streams; ++x)
    {
        ICP_CUDA(cudaEventRecord(b->ev_done[x], b->streams[x]));
        ICP_CUDA(cudaStreamWaitEvent(st, b->ev_done[x], 0));
    }
    return ICP_OK;
}

extern "C" int icp_batch_set_slices(icp_batch *b, uint32_t n_slices)
{ ICP_ENTER_OBJ(b);
    if (!b || n_slices == 0 || n_slices > 256) { icp_set_error("icp_batch_set_slices: 1..256"); return ICP_ERR_ARG; }
    b->n_slices = n_slices;
    return ICP_OK;
}

// Host-buffer entry of the batch (what a caller holding frames in host memory uses).  Blocking; the 8-float poses of
// all pairs are returned in h_T8.  Results are identical to icp_batch_upload + icp_batch_register + icp_batch_read_poses.
extern "C" int icp_batch_register_host(icp_batch *b, const float *h_F, const float *h_M, uint32_t n_iters, uint32_t n_slices, float *h_T8)
{ ICP_ENTER_OBJ(b);
    if (!b || !h_F || !h_M || n_iters == 0) { icp_set_error("icp_batch_register_host: bad argument"); return ICP_ERR_ARG; }
    if (n_slices == 0) n_slices = b->n_slices;
    if (b->pending) { icp_set_error("icp_batch_register_host: an asynchronous registration is pending (icp_batch_collect first)"); return ICP_ERR_ARG; }
    ICP_CHECK(batch_run_sliced(b, n_iters, n_slices, h_F, h_M, nullptr));
    cudaStream_t st = b->ctx->stream;
    ICP_CUDA(cudaMemcpyAsync(b->h_T, b->T, (size_t)b->n_pairs * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    ICP_CUDA(cudaStreamSynchronize(st));
    if (h_T8) memcpy(h_T8, b->h_T, (size_t)b->n_pairs * 8 * sizeof(float));
    return ICP_OK;
}

// Asynchronous pair of the host-buffer entry, for callers that stream batches: enqueue (uploads, registration, d2h of the
// poses into the batch's pinned staging buffer) on the batch's OWN streams and return; icp_batch_collect waits and hands
// the poses out.  Two icp_batch objects used alternately overlap the uploads of one step with the registration of the
// previous one (bench.py's e2e leg).  Ordered only with respect to the same batch: the host buffers must stay valid and
// the batch untouched until icp_batch_collect returns.
extern "C" int icp_batch_register_host_async(icp_batch *b, const float *h_F, const float *h_M, uint32_t n_iters, uint32_t n_slices)
{ ICP_ENTER_OBJ(b);
    if (!b || !h_F || !h_M || n_iters == 0) { icp_set_error("icp_batch_register_host_async: bad argument"); return ICP_ERR_ARG; }
    if (b->pending) { icp_set_error("icp_batch_register_host_async: a registration is already pending (icp_batch_collect first)"); return ICP_ERR_ARG; }
    if (n_slices == 0) n_slices = b->n_slices;
    if (!b->home_stream) ICP_CUDA(cudaStreamCreateWithFlags(&b->home_stream, cudaStreamNonBlocking));
    ICP_CHECK(batch_run_sliced(b, n_iters, n_slices, h_F, h_M, b->home_stream));
    ICP_CUDA(cudaMemcpyAsync(b->h_T, b->T, (size_t)b->n_pairs * 8 * sizeof(float), cudaMemcpyDeviceToHost, b->home_stream));
    b->pending = true;
    return ICP_OK;
}

extern "C" int icp_batch_collect(icp_batch *b, float *h_T8)
{ ICP_ENTER_OBJ(b);
    if (!b || !b->pending) { icp_set_error("icp_batch_collect: nothing pending"); return ICP_ERR_ARG; }
    ICP_CUDA(cudaStreamSynchronize(b->home_stream));
    b->pending = false;
    if (h_T8) memcpy(h_T8, b->h_T, (size_t)b->n_pairs * 8 * sizeof(float));
    return ICP_OK;
}

extern "C" int icp_batch_read_poses(icp_batch *b, float *h_T8, float *h_T16)
{ ICP_ENTER_OBJ(b);
    cudaStream_t st = b->ctx->stream;
    ICP_CUDA(cudaMemcpyAsync(b->h_T, b->T, (size_t)b->n_pairs * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (h_T16) ICP_CUDA(cudaMemcpyAsync(b->h_state, b->state, (size_t)b->n_pairs * sizeof(icp_state), cudaMemcpyDeviceToHost, st));
    ICP_CUDA(cudaStreamSynchronize(st));
    if (h_T8) memcpy(h_T8, b->h_T, (size_t)b->n_pairs * 8 * sizeof(float));
    if (h_T16)
        for (uint32_t p = 0; p < b->n_pairs; ++p)
        {
            const icp_state *s = b->h_state + p;
            float *o = h_T16 + (size_t)p * 16;
            for (int i = 0; i < 3; ++i)
            {
                for (int j = 0; j < 3; ++j) o[i * 4 + j] = s->s * s->R[i * 3 + j];
                o[i * 4 + 3] = s->t[i];
            }
            o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
        }
    return ICP_OK;
}

extern "C" void *icp_batch_debug_ptr(icp_batch *b, const char *name)
{
    // "<field>" of pair 0, or "<field>@<pair>"
    char field[32];
    uint32_t pair = 0;
    const char *at = strchr(name, '@');
    size_t len = at ? (size_t)(at - name) : strlen(name);
    if (len >= sizeof(field)) return nullptr;
    memcpy(field, name, len); field[len] = 0;
    if (at) pair = (uint32_t)atoi(at + 1);
    if (pair >= b->n_pairs) return nullptr;
    const PairPtrs &P = b->h_table[pair];
#define NAME(n, p) if (!strcmp(field, n)) return (void *)(p)
    NAME("gt", b->gt + (size_t)pair * 8); NAME("T", P.T); NAME("state", P.state);
    NAME("reps", P.reps); NAME("rep_id", P.rep_id); NAME("N", P.N); NAME("O", P.O); NAME("perm", P.perm); NAME("Xp", P.Xp);
    NAME("q_rep", P.q_rep); NAME("qperm", P.qperm); NAME("Nq", P.Nq); NAME("Oq", P.Oq); NAME("NN_ID", P.NNID);
    NAME("W", P.W); NAME("sum_w", P.sum_w); NAME("mean", P.mean); NAME("S", P.S); NAME("Tk", P.Tk);
    NAME("fxyz", P.fxyz); NAME("mxyz", P.mxyz); NAME("F", P.F); NAME("M", P.M); NAME("evals", P.evals); NAME("nnd", P.nnd); NAME("nn_o", P.nn_o); NAME("nbx", P.nbx); NAME("prof", P.prof); NAME("lrank", P.lrank); NAME("wconst", P.wconst);
#undef NAME
    return nullptr;
}

// time one of the fused kernels standalone on the batch's current data (roofline measurement):
// which: 0 = A (assign/search side), 1 = B (colscan), 2 = C (search), 3 = D (reduce+solve; advances the poses).
// Returns the average device time per launch in ms (CUDA events on the context stream).
int fused_launch_one(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, int which);

extern "C" int icp_batch_time_kernel(icp_batch *b, int which, uint32_t n_launches, float *ms_avg)
{ ICP_ENTER_OBJ(b);
    if (n_launches == 0 || which < 0 || which > 3) { icp_set_error("icp_batch_time_kernel: bad argument"); return ICP_ERR_ARG; }
    cudaStream_t st = b->ctx->stream;
    // The work of an iteration depends on where the registration stands (kernel A's pruning, the settle test of kernel C'),
    // and B / D are not idempotent: the timed launches are the FIRST n_launches iterations of a fresh registration
    // (reset + buildRBC, then A, B, C, D in stream order, unfused), the events bracket only the kernel of interest.
    // Pass 0 is a warm-up; the result is the average over pass 1.
    double total = 0.0;
    for (int pass = 0; pass < 2; ++pass)
    {
        k_batch_reset<<<div_up(b->n_pairs, 128), 128, 0, st>>>(b->state, b->T, b->loop, b->n_pairs, (int32_t)n_launches);
        ICP_CHECK(fused_launch_build(st, b->cfg, b->table, b->n_pairs, b->lm_w, b->lm_h));
        for (uint32_t i = 0; i < n_launches; ++i)
        {
            for (int k = 0; k < 4; ++k)
            {
                if (k == which) ICP_CUDA(cudaEventRecord(b->ctx->ev0, st));
                ICP_CHECK(fused_launch_one(st, b->cfg, b->table, b->n_pairs, k));
                if (k == which) ICP_CUDA(cudaEventRecord(b->ctx->ev1, st));
            }
            ICP_CUDA(cudaEventSynchronize(b->ctx->ev1));
            ICP_CUDA(cudaStreamSynchronize(st));
            float ms = 0.f;
            ICP_CUDA(cudaEventElapsedTime(&ms, b->ctx->ev0, b->ctx->ev1));
            if (pass == 1) total += ms;
        }
    }
    *ms_avg = (float)(total / n_launches);
    return ICP_OK;
}

extern "C" uint32_t icp_batch_slices(icp_batch *b) { ICP_ENTER_OBJ(b); return b ? b->n_slices : 0u; }
// kernel-C flavour of the batch: 0 = k_search<L>, 1 = k_search_grouped, 2 = k_colscan_sort + k_search_sorted, 3 = sorted records + k_search_span
extern "C" int icp_batch_cmode(icp_batch *b) { ICP_ENTER_OBJ(b); return b ? b->cfg.Cmode : -1; }

extern "C" int icp_batch_config(icp_batch *b, uint32_t *QB, uint32_t *nbA, int *S, int *CL, int *L)
{ ICP_ENTER_OBJ(b);
    if (QB) *QB = b->cfg.QB;
    if (nbA) *nbA = b->cfg.nbA;
    if (S) *S = b->cfg.S;
    if (CL) *CL = b->cfg.CL;
    if (L) *L = b->cfg.L;
    return ICP_OK;
}
