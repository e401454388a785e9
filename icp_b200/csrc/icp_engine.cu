// icp_engine.cu -- ICPStep<CR,CW> / ICP<CR,CW> replacement: device-resident registration loop.
//
// The reference enqueues >= 17 kernels per iteration, blocks on a 32-byte read, does the pose update with
// Eigen on the host and writes it back (algorithms.cpp:4670-4698).  Here the pose update runs on the device
// and the whole loop is one CUDA graph: either an unrolled chain of n iterations or a conditional WHILE node
// whose body is one iteration and whose condition is ICP::check() evaluated by the last kernel
// (cudaGraphSetConditional) -- no host round trip per iteration.
#include "icp_engine.cuh"
#include "icp_solve.cuh"
#include <string.h>
#include <stdlib.h>

// ------------------------------------------------------------------------------------------------
// last kernel of an iteration: solve is done, accumulate pose + loop control
// ------------------------------------------------------------------------------------------------
__global__ void k_step_finish(DevState *state, const float *__restrict__ Tk, const float *__restrict__ Rk, float *__restrict__ T,
                              LoopParams *lp, cudaGraphConditionalHandle handle, int use_handle)
{
    if (threadIdx.x != 0) return;
    float tk[8], rk[9], t8[8];
    for (int i = 0; i < 8; ++i) tk[i] = Tk[i];
    if (Rk) for (int i = 0; i < 9; ++i) rk[i] = Rk[i];
    solve::accumulate(state, tk, Rk ? rk : nullptr, t8);
    for (int i = 0; i < 8; ++i) T[i] = t8[i];
    unsigned cont;
    const int left = lp->iters_left - 1;
    lp->iters_left = left;
    if (lp->check)
    {
        solve::check_convergence(state, lp->max_iterations, lp->angle_thr, lp->trans_thr);
        cont = (state->done == 0u && left > 0) ? 1u : 0u;
    }
    else
    {
        state->k = state->k + 1;
        cont = left > 0 ? 1u : 0u;
    }
    if (use_handle) cudaGraphSetConditional(handle, cont);
}

__global__ void k_state_reset(DevState *state, float *T, int reset_T)
{
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 9; ++i) { state->R[i] = (i % 4 == 0) ? 1.f : 0.f; state->Rk[i] = (i % 4 == 0) ? 1.f : 0.f; }
    for (int i = 0; i < 4; ++i) { state->q[i] = (i == 3) ? 1.f : 0.f; state->qk[i] = (i == 3) ? 1.f : 0.f; }
    for (int i = 0; i < 3; ++i) { state->t[i] = 0.f; state->tk[i] = 0.f; }
    state->s = 1.f; state->sk = 1.f;
    state->k = 0; state->done = 0;
    if (reset_T) { T[0] = 0.f; T[1] = 0.f; T[2] = 0.f; T[3] = 1.f; T[4] = 0.f; T[5] = 0.f; T[6] = 0.f; T[7] = 1.f; }
}

__global__ void k_count_e1(unsigned long long *evals, unsigned long long add)
{
    if (threadIdx.x == 0) evals[0] += add;
}

// ------------------------------------------------------------------------------------------------
// construction / init
// ------------------------------------------------------------------------------------------------
extern "C" int icp_step_create(icp_ctx *ctx, int rot_cfg, int w_cfg, icp_step **out)
{ ICP_ENTER(ctx);
    if (!ctx || !out) { icp_set_error("icp_step_create: null argument"); return ICP_ERR_ARG; }
    if (rot_cfg != ICP_ROT_EIGEN && rot_cfg != ICP_ROT_POWER_METHOD) { icp_set_error("icp_step_create: bad rot_cfg"); return ICP_ERR_ARG; }
    if (w_cfg != ICP_W_REGULAR && w_cfg != ICP_W_WEIGHTED) { icp_set_error("icp_step_create: bad w_cfg"); return ICP_ERR_ARG; }
    icp_step *s = new icp_step();
    s->ctx = ctx; s->rot_cfg = rot_cfg; s->w_cfg = w_cfg;
    *out = s;
    return ICP_OK;
}

void engine_drop_graphs(icp_step *s)
{
    for (auto &kv : s->unrolled) cudaGraphExecDestroy(kv.second);
    s->unrolled.clear();
    if (s->while_exec) { cudaGraphExecDestroy(s->while_exec); s->while_exec = nullptr; }
    if (s->build_exec) { cudaGraphExecDestroy(s->build_exec); s->build_exec = nullptr; }
    s->while_failed = false;
}

extern "C" void icp_step_destroy(icp_step *s)
{
    if (!s) return;
    IcpDeviceGuard guard__(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    engine_drop_graphs(s);
    if (s->arena) cudaFree(s->arena);
    if (s->own_F && s->F) cudaFree(s->F);
    if (s->own_M && s->M) cudaFree(s->M);
    if (s->own_T && s->T) cudaFree(s->T);
    if (s->h_loop) cudaFreeHost(s->h_loop);
    if (s->h_state) cudaFreeHost(s->h_state);
    delete s;
}

extern "C" int icp_step_bind(icp_step *s, int mem, void *d_ptr)
{ ICP_ENTER_OBJ(s);
    // takes effect at the next init() (algorithms.cpp:216-221: init only creates what is still null)
    // point sets move as 256-bit requests in the fused kernels: 32-byte alignment (any cudaMalloc / icp_malloc pointer has 256)
    if ((mem == ICP_MEM_D_IN_F || mem == ICP_MEM_D_IN_M) && ((uintptr_t)d_ptr & 31u)) { icp_set_error("icp_step_bind: point buffers must be 32-byte aligned"); return ICP_ERR_ARG; }
    cudaStreamSynchronize(s->ctx->stream);
    s->inited = false;
    switch (mem)
    {
        case ICP_MEM_D_IN_F: if (s->own_F && s->F) cudaFree(s->F); s->F = (float *)d_ptr; s->own_F = false; break;
        case ICP_MEM_D_IN_M: if (s->own_M && s->M) cudaFree(s->M); s->M = (float *)d_ptr; s->own_M = false; break;
        case ICP_MEM_D_IO_T: if (s->own_T && s->T) cudaFree(s->T); s->T = (float *)d_ptr; s->own_T = false; break;
        default: icp_set_error("icp_step_bind: unknown memory id %d", mem); return ICP_ERR_ARG;
    }
    return ICP_OK;
}

static const char *step_class_name(const icp_step *s)
{
    if (s->rot_cfg == ICP_ROT_POWER_METHOD)
        return s->w_cfg ? "ICPStep<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::WEIGHTED>" : "ICPStep<ICPStepConfigT::POWER_METHOD, ICPStepConfigW::REGULAR>";
    return s->w_cfg ? "ICPStep<ICPStepConfigT::EIGEN, ICPStepConfigW::WEIGHTED>" : "ICPStep<ICPStepConfigT::EIGEN, ICPStepConfigW::REGULAR>";
}

static size_t carve_all(icp_step *s, void *base)
{
    Carver cv(base);
    const uint32_t m = s->m, nr = s->nr;
    s->reps = cv.take<float>((size_t)nr * 8);
    s->Xp = cv.take<float>((size_t)m * 8);
    s->rep_id = cv.take<uint32_t>(m);
    s->N = cv.take<uint32_t>(nr);
    s->O = cv.take<uint32_t>(nr);
    s->perm = cv.take<uint32_t>(m);
    s->Mt = cv.take<float>((size_t)m * 8);
    s->Qp = cv.take<float>((size_t)m * 8);
    s->NN = cv.take<float>((size_t)m * 8);
    s->W = cv.take<float>(m);
    s->mean = cv.take<float>(8);
    s->DF = cv.take<float>((size_t)m * 4);
    s->DM = cv.take<float>((size_t)m * 4);
    s->S = cv.take<float>(16);
    s->Tk = cv.take<float>(8);
    s->Rk = cv.take<float>(12);
    s->q_rep = cv.take<uint32_t>(m);
    s->qperm = cv.take<uint32_t>(m);
    s->Nq = cv.take<uint32_t>(nr);
    s->Oq = cv.take<uint32_t>(nr);
    s->NNID = cv.take<icp_dist_id>(m);
    s->sum_w = cv.take<double>(2);
    s->state = cv.take<DevState>(1);
    s->loop = cv.take<LoopParams>(1);
    s->evals = cv.take<unsigned long long>(4);
    s->sort_scr = cv.take<char>(SortScratch::bytes(m, nr));
    const size_t e = reduce_scratch_elems(m);
    s->red_f = cv.take<float>(e + 8);
    s->red_d = cv.take<double>(2 * e);
    s->mean_scr = cv.take<float>(16 * e);
    s->sij_part = cv.take<float>(sij_partials_elems(m));
    s->sij_scr = cv.take<float>(sij_scratch_elems(m));
    s->fused = cv.take<char>(fused_workspace_bytes(m, nr, s->ctx->sm_count));
    return cv.off + 256;
}

extern "C" int icp_step_init(icp_step *s, uint32_t m, uint32_t nr, float alpha, float c, uint32_t lm_w, uint32_t lm_h)
{ ICP_ENTER_OBJ(s);
    const char *cls = step_class_name(s);
    if (m == 0) ICP_CONFIG_FAIL(cls, "The sets of landmarks cannot have zero points");              // algorithms.cpp:4413
    if (nr == 0) ICP_CONFIG_FAIL(cls, "The sets of representatives cannot have zero points");       // :4416
    if (alpha == 0.f) ICP_CONFIG_FAIL(cls, "The alpha parameter cannot be equal to zero");          // :4419
    if (m > (1u << 20)) ICP_CONFIG_FAIL(cls, "The current configuration supports sets of up to 1048576 points");
    if (nr > 4096) ICP_CONFIG_FAIL(cls, "The current configuration supports up to 4096 representatives");
    if (lm_w == 0 && lm_h == 0) { lm_w = 128; lm_h = 128; }
    if ((uint64_t)lm_w * lm_h != m) ICP_CONFIG_FAIL(cls, "The landmark grid (lm_w x lm_h) must hold exactly m points");
    if (nr % 4) ICP_CONFIG_FAIL("ICPReps", "The number of representatives has to be a multiple of 4");
    if (m % 2) ICP_CONFIG_FAIL("ICPWeights", "The number of elements in the array must be a multiple of 2");
    {
        uint32_t nrx, nry;
        icp_rep_grid(nr, &nrx, &nry);
        if (nrx * nry != nr) ICP_CONFIG_FAIL("ICPReps", "The number of representatives has to be a power of 2");
        if (lm_w % nrx || lm_h % nry || lm_w / nrx < 2 || lm_h / nry < 2)
            ICP_CONFIG_FAIL("ICPReps", "The landmark grid is not divisible into the representative grid");
    }
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    engine_drop_graphs(s);
    if (s->arena) { ICP_CUDA(cudaFree(s->arena)); s->arena = nullptr; }
    s->m = m; s->nr = nr; s->a = alpha; s->c = c; s->lm_w = lm_w; s->lm_h = lm_h;
    if (!s->metric_override) icp_metric_weights(alpha, &s->fg, &s->fp);
    const size_t fm_bytes = (size_t)m * 8 * sizeof(float);
    if (!s->F) { ICP_CUDA(cudaMalloc((void **)&s->F, fm_bytes)); s->own_F = true; }
    if (!s->M) { ICP_CUDA(cudaMalloc((void **)&s->M, fm_bytes)); s->own_M = true; }
    if (!s->T) { ICP_CUDA(cudaMalloc((void **)&s->T, 8 * sizeof(float))); s->own_T = true; }
    s->arena_bytes = carve_all(s, nullptr);
    ICP_CUDA(cudaMalloc(&s->arena, s->arena_bytes));
    ICP_CUDA(cudaMemsetAsync(s->arena, 0, s->arena_bytes, s->ctx->stream));
    carve_all(s, s->arena);
    if (!s->h_loop) ICP_CUDA(cudaMallocHost((void **)&s->h_loop, sizeof(LoopParams)));
    if (!s->h_state) ICP_CUDA(cudaMallocHost((void **)&s->h_state, sizeof(icp_state)));
    s->inited = true;
    s->rbc_mode = -1;
    ICP_CHECK(fused_prepare(s));
    return icp_step_reset(s);
}

extern "C" void *icp_step_buffer(icp_step *s, int mem)
{
    switch (mem)
    {
        case ICP_MEM_D_IN_F: return s->F;
        case ICP_MEM_D_IN_M: return s->M;
        case ICP_MEM_D_IO_T: return s->T;
        default: return nullptr;
    }
}

extern "C" int icp_step_write(icp_step *s, int mem, const void *h_src, int block)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_write: init() first"); return ICP_ERR_ARG; }
    void *dst = icp_step_buffer(s, mem);
    if (!dst) { icp_set_error("icp_step_write: unknown memory id %d", mem); return ICP_ERR_ARG; }
    const size_t bytes = (mem == ICP_MEM_D_IO_T) ? 8 * sizeof(float) : (size_t)s->m * 8 * sizeof(float);
    return icp_memcpy_h2d(s->ctx, dst, h_src, bytes, block);
}

extern "C" int icp_step_reset(icp_step *s)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_reset: init() first"); return ICP_ERR_ARG; }
    k_state_reset<<<1, 32, 0, s->ctx->stream>>>(s->state, s->T, 1);
    ICP_LAUNCH_CHECK();
    ICP_CUDA(cudaMemsetAsync(s->evals, 0, 4 * sizeof(unsigned long long), s->ctx->stream));
    return ICP_OK;
}

extern "C" int icp_step_set_alpha(icp_step *s, float alpha)
{ ICP_ENTER_OBJ(s);
    if (alpha == 0.f) ICP_CONFIG_FAIL(step_class_name(s), "The alpha parameter cannot be equal to zero");
    s->a = alpha;
    s->metric_override = false;
    icp_metric_weights(alpha, &s->fg, &s->fp);
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    engine_drop_graphs(s);      // kernel arguments are baked into the graphs
    return ICP_OK;
}
extern "C" int icp_step_set_scaling(icp_step *s, float c)
{ ICP_ENTER_OBJ(s);
    s->c = c;
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    engine_drop_graphs(s);
    return ICP_OK;
}
extern "C" int icp_step_set_metric(icp_step *s, float f_g, float f_p)
{ ICP_ENTER_OBJ(s);
    s->fg = f_g; s->fp = f_p; s->metric_override = true;
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    engine_drop_graphs(s);
    return ICP_OK;
}
extern "C" int icp_step_set_mode(icp_step *s, int mode)
{ ICP_ENTER_OBJ(s);
    if (mode != ICP_MODE_STAGED && mode != ICP_MODE_FUSED) { icp_set_error("icp_step_set_mode: bad mode"); return ICP_ERR_ARG; }
    if (mode != s->mode)
    {
        ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
        engine_drop_graphs(s);
        s->mode = mode;
        // the lane order kernel A keeps from one fused iteration to the next is stale after staged iterations
        if (s->inited && mode == ICP_MODE_FUSED) ICP_CHECK(fused_invalidate(s, s->ctx->stream, true, true));
    }
    return ICP_OK;
}
extern "C" int icp_step_set_count_evals(icp_step *s, int on)
{ ICP_ENTER_OBJ(s);
    if ((on != 0) != s->count_evals)
    {
        ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
        engine_drop_graphs(s);
        s->count_evals = on != 0;
        if (s->inited) ICP_CHECK(fused_prepare(s));
    }
    return ICP_OK;
}

// ------------------------------------------------------------------------------------------------
// buildRBC (algorithms.cpp:4655-4660): representatives + RBC construct over the fixed set; k = 0
// ------------------------------------------------------------------------------------------------
int engine_enqueue_build(icp_step *s, cudaStream_t st)
{
    ICP_CHECK(launch_get_reps(st, s->F, s->lm_w, s->lm_h, s->nr, s->reps));
    ICP_CHECK(launch_nearest_rep(st, s->ctx->sm_count, s->F, s->m, s->reps, s->nr, s->fg, s->fp, nullptr, nullptr, s->rep_id));
    SortScratch sc; sc.carve(s->sort_scr, s->m, s->nr);
    ICP_CHECK(launch_counting_sort(st, s->rep_id, s->m, s->nr, sc, s->N, s->O, s->perm, s->F, s->Xp));
    return ICP_OK;
}

__global__ void k_reset_k(DevState *state) { if (threadIdx.x == 0) { state->k = 0; state->done = 0; } }

extern "C" int icp_step_build_rbc(icp_step *s)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_build_rbc: init() first"); return ICP_ERR_ARG; }
    cudaStream_t st = s->ctx->stream;
    if (s->mode == ICP_MODE_FUSED) ICP_CHECK(fused_enqueue_build(s, st));
    else ICP_CHECK(engine_enqueue_build(s, st));
    s->rbc_mode = s->mode;
    k_reset_k<<<1, 32, 0, st>>>(s->state);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// Called at the top of every run entry.  (1) A fused iteration after a STAGED buildRBC: the staged build wrote the RBC itself
// (reps, rep_id, N, O, perm, X_p) but not the fused kernels' acceleration tables (representative neighbour rows, constant-lane
// flags, temporal bounds), which would still describe the previous fixed set -- rebuild them with the fused build (same RBC
// bit for bit).  (2) The caller may have rewritten the moving set since the last run call (the reference allows it without a
// new buildRBC: the RBC only depends on F), so the temporal-pruning bounds are not trusted across run calls.
static int engine_prepare_run(icp_step *s)
{
    if (s->mode != ICP_MODE_FUSED) return ICP_OK;
    if (s->rbc_mode == ICP_MODE_STAGED)
    {
        ICP_CHECK(fused_enqueue_build(s, s->ctx->stream));
        s->rbc_mode = ICP_MODE_FUSED;
    }
    return fused_invalidate(s, s->ctx->stream, false, true);
}

// ------------------------------------------------------------------------------------------------
// one iteration, staged: one kernel (family) per reference stage, same order as algorithms.cpp:4673-4679
// ------------------------------------------------------------------------------------------------
static int staged_enqueue_iteration(icp_step *s, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle, cudaEvent_t *marks)
{
    const uint32_t m = s->m, nr = s->nr;
    int mk = 0;
#define MARK() do { if (marks) ICP_CUDA(cudaEventRecord(marks[mk++], st)); } while (0)
    MARK();
    ICP_CHECK(launch_transform_q(st, s->M, s->T, s->Mt, m));                                          // transform.run ()
    MARK();
    ICP_CHECK(launch_nearest_rep(st, s->ctx->sm_count, s->Mt, m, s->reps, nr, s->fg, s->fp, nullptr, nullptr, s->q_rep));   // rbcS.run ()
    SortScratch sc; sc.carve(s->sort_scr, m, nr);
    ICP_CHECK(launch_counting_sort(st, s->q_rep, m, nr, sc, s->Nq, s->Oq, s->qperm, nullptr, nullptr));
    ICP_CHECK(launch_rbc_stage2(st, s->Mt, s->q_rep, s->qperm, m, s->Xp, m, s->O, s->N, s->fg, s->fp, s->Qp, s->NN, s->NNID,
                                s->count_evals ? s->evals + 1 : nullptr));
    if (s->count_evals) { k_count_e1<<<1, 32, 0, st>>>(s->evals, (unsigned long long)m * nr); ICP_LAUNCH_CHECK(); }
    MARK();
    if (s->w_cfg)
    {
        ICP_CHECK(launch_weights(st, s->NNID, s->W, s->sum_w, m, s->red_f, s->red_d));                // weights.run ()
        MARK();
        ICP_CHECK(launch_mean(st, s->NN, s->Qp, s->W, s->sum_w, s->mean, m, s->mean_scr));             // means.run ()
    }
    else
    {
        MARK();
        ICP_CHECK(launch_mean(st, s->NN, s->Qp, nullptr, nullptr, s->mean, m, s->mean_scr));
    }
    MARK();
    ICP_CHECK(launch_devs(st, s->NN, s->Qp, s->mean, s->DF, s->DM, m));                                // devs.run ()
    MARK();
    ICP_CHECK(launch_sij(st, s->DM, s->DF, s->w_cfg ? s->W : nullptr, s->S, m, s->c, s->sij_part, s->sij_scr));   // matrixS.run ()
    MARK();
    if (s->rot_cfg == ICP_ROT_POWER_METHOD) ICP_CHECK(launch_power_method(st, s->S, s->mean, s->Tk));  // powMethod.run ()
    else ICP_CHECK(launch_svd_solve(st, s->S, s->mean, s->Tk, s->Rk));                                 // host Eigen SVD in the reference
    k_step_finish<<<1, 32, 0, st>>>(s->state, s->Tk, s->rot_cfg == ICP_ROT_POWER_METHOD ? nullptr : s->Rk, s->T, s->loop, handle, use_handle);
    ICP_LAUNCH_CHECK();
    MARK();
#undef MARK
    return ICP_OK;
}

int engine_enqueue_iteration(icp_step *s, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle)
{
    if (s->mode == ICP_MODE_FUSED) return fused_enqueue_iteration(s, st, handle, use_handle);
    return staged_enqueue_iteration(s, st, handle, use_handle, nullptr);
}

// ------------------------------------------------------------------------------------------------
// graphs
// ------------------------------------------------------------------------------------------------
__global__ void k_set_loop(LoopParams *lp, DevState *state, LoopParams v)
{
    if (threadIdx.x == 0) { *lp = v; state->done = 0; }
}

// loop control travels as a kernel argument (by value): no host buffer lifetime, no synchronisation
static int set_loop_params(icp_step *s, uint32_t check, uint32_t max_it, int32_t iters, double ang, double tr)
{
    LoopParams v;
    v.check = check; v.max_iterations = max_it; v.iters_left = iters; v.angle_thr = ang; v.trans_thr = tr; v.pad = 0;
    k_set_loop<<<1, 32, 0, s->ctx->stream>>>(s->loop, s->state, v);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

static int get_unrolled(icp_step *s, uint32_t n, cudaGraphExec_t *out)
{
    auto it = s->unrolled.find(n);
    if (it != s->unrolled.end()) { *out = it->second; return ICP_OK; }
    cudaStream_t st = s->ctx->stream;
    cudaGraph_t g = nullptr;
    ICP_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = ICP_OK;
    for (uint32_t i = 0; i < n && rc == ICP_OK; ++i) rc = engine_enqueue_iteration(s, st, 0, 0);
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (rc != ICP_OK) { if (g) cudaGraphDestroy(g); return rc; }
    ICP_CUDA(e);
    cudaGraphExec_t ex = nullptr;
    ICP_CUDA(cudaGraphInstantiate(&ex, g, 0));
    cudaGraphDestroy(g);
    if (s->unrolled.size() >= 16)       // bounded cache: a caller sweeping iteration counts does not accumulate graphs
    {
        ICP_CUDA(cudaStreamSynchronize(st));
        for (auto &kv : s->unrolled) cudaGraphExecDestroy(kv.second);
        s->unrolled.clear();
    }
    s->unrolled[n] = ex;
    *out = ex;
    return ICP_OK;
}

// graph = [conditional WHILE node { one iteration; last kernel sets the condition }]
static int get_while(icp_step *s, cudaGraphExec_t *out)
{
    if (s->while_exec) { *out = s->while_exec; return ICP_OK; }
    if (s->while_failed) return ICP_ERR_CUDA;
    cudaStream_t st = s->ctx->stream;
    cudaGraph_t g = nullptr;
    cudaError_t e;
    cudaGraphConditionalHandle handle;
    cudaGraphNode_t node;
    cudaGraphNodeParams p = {};
    bool capturing = false;
    do
    {
        if ((e = cudaGraphCreate(&g, 0)) != cudaSuccess) break;
        if ((e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault)) != cudaSuccess) break;
        p.type = cudaGraphNodeTypeConditional;
        p.conditional.handle = handle;
        p.conditional.type = cudaGraphCondTypeWhile;
        p.conditional.size = 1;
        if ((e = cudaGraphAddNode(&node, g, nullptr, 0, &p)) != cudaSuccess) break;
        cudaGraph_t body = p.conditional.phGraph_out[0];
        if ((e = cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) break;
        capturing = true;
        int rc = engine_enqueue_iteration(s, st, handle, 1);
        cudaGraph_t dummy = nullptr;
        e = cudaStreamEndCapture(st, &dummy);
        capturing = false;
        if (rc != ICP_OK) { e = cudaErrorUnknown; break; }
        if (e != cudaSuccess) break;
        cudaGraphExec_t ex = nullptr;
        if ((e = cudaGraphInstantiate(&ex, g, 0)) != cudaSuccess) break;
        s->while_exec = ex;
    } while (0);
    if (capturing) { cudaGraph_t dummy = nullptr; cudaStreamEndCapture(st, &dummy); }
    if (g) cudaGraphDestroy(g);
    if (!s->while_exec)
    {
        s->while_failed = true;
        cudaGetLastError();
        icp_set_error("conditional WHILE graph unavailable: %s", cudaGetErrorString(e));
        return ICP_ERR_CUDA;
    }
    *out = s->while_exec;
    return ICP_OK;
}

// Which loop engine runs a fused registration: "graph" (unrolled / conditional-WHILE CUDA graphs of the four fused kernels) or
// "persistent" (one cooperative launch, icp_fused.cu: k_icp_persistent).  Picked by measurement (profiles/r02_latency_engines.md);
// ICP_B200_ENGINE overrides it.
// Measured on one B200, 16384 / 256, us per ICP iteration (power method / SVD): unrolled graph 44.5 / 41.8, persistent kernel
// 44.3 / 42.0, conditional WHILE graph 49.8 / 47.2, plain stream launches 45.5 / 42.9.  So a fixed iteration count replays
// the unrolled graph (cached per count), and whatever needs a device-side loop -- the thresholded ICP::run (), very long
// runs -- takes the persistent kernel instead of the WHILE graph when the size is eligible.
static bool engine_prefers_persistent(const icp_step *s, bool device_loop)
{
    if (s->mode != ICP_MODE_FUSED) return false;
    if (const char *e = getenv("ICP_B200_ENGINE")) return e[0] == 'p' || e[0] == 'P';
    return device_loop;
}
static int read_state(icp_step *s);

// n_iters x ICPStep::run
extern "C" int icp_step_run(icp_step *s, uint32_t n_iters)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_run: init() first"); return ICP_ERR_ARG; }
    if (n_iters == 0) return ICP_OK;
    ICP_CHECK(engine_prepare_run(s));
    ICP_CHECK(set_loop_params(s, 0, 0, (int32_t)n_iters, 0.0, 0.0));
    if (engine_prefers_persistent(s, n_iters > 128))
    {
        int ok = 0;
        ICP_CHECK(fused_enqueue_persistent(s, s->ctx->stream, n_iters, &ok));
        if (ok) return ICP_OK;
    }
    cudaGraphExec_t ex = nullptr;
    // a fixed iteration count replays an unrolled graph (cached per count; measured ~5 us/iteration cheaper than a
    // conditional WHILE node); very long runs and the thresholded ICP::run use the WHILE graph
    if (n_iters > 128 && get_while(s, &ex) == ICP_OK)
    {
        ICP_CUDA(cudaGraphLaunch(ex, s->ctx->stream));
        return ICP_OK;
    }
    ICP_CHECK(get_unrolled(s, n_iters, &ex));
    ICP_CUDA(cudaGraphLaunch(ex, s->ctx->stream));
    return ICP_OK;
}

// variants for measurement: 0 = plain stream launches, 1 = unrolled graph, 2 = conditional WHILE graph
extern "C" int icp_step_run_variant(icp_step *s, uint32_t n_iters, int variant)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_run_variant: init() first"); return ICP_ERR_ARG; }
    if (n_iters == 0) return ICP_OK;
    ICP_CHECK(engine_prepare_run(s));
    ICP_CHECK(set_loop_params(s, 0, 0, (int32_t)n_iters, 0.0, 0.0));
    cudaGraphExec_t ex = nullptr;
    if (variant == 0)
    {
        for (uint32_t i = 0; i < n_iters; ++i) ICP_CHECK(engine_enqueue_iteration(s, s->ctx->stream, 0, 0));
        return ICP_OK;
    }
    if (variant == 3)
    {
        // persistent cooperative kernel: one launch, software grid barriers between the phases (fused mode, eligible sizes)
        int ok = 0;
        if (s->mode == ICP_MODE_FUSED) ICP_CHECK(fused_enqueue_persistent(s, s->ctx->stream, n_iters, &ok));
        if (!ok) { icp_set_error("icp_step_run_variant: the persistent engine is not available for this mode / size / device"); return ICP_ERR_ARG; }
        return ICP_OK;
    }
    if (variant == 1) ICP_CHECK(get_unrolled(s, n_iters, &ex));
    else ICP_CHECK(get_while(s, &ex));
    ICP_CUDA(cudaGraphLaunch(ex, s->ctx->stream));
    return ICP_OK;
}

static int read_state(icp_step *s)
{
    ICP_CUDA(cudaMemcpyAsync(s->h_state, s->state, sizeof(icp_state), cudaMemcpyDeviceToHost, s->ctx->stream));
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    return ICP_OK;
}

// ICP::run: first step unconditionally, then while (check ()) step  (algorithms.cpp:4807-4814)
extern "C" int icp_run(icp_step *s, uint32_t max_iterations, double angle_threshold_deg, double translation_threshold_mm, uint32_t *k_out)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_run: init() first"); return ICP_ERR_ARG; }
    ICP_CHECK(engine_prepare_run(s));
    // ICP::check() stops at k == max_iterations (algorithms.cpp:4828); max_iterations = 0 never matches, i.e. "until the
    // thresholds are met" in the reference.  The device loop additionally carries a safety bound of 2^31 - 1 steps.
    const int32_t bound = 0x7fffffff;
    ICP_CHECK(set_loop_params(s, 1, max_iterations, bound, angle_threshold_deg, translation_threshold_mm));
    cudaGraphExec_t ex = nullptr;
    int persistent_ok = 0;
    if (engine_prefers_persistent(s, true)) ICP_CHECK(fused_enqueue_persistent(s, s->ctx->stream, (uint32_t)bound, &persistent_ok));
    if (persistent_ok) ICP_CHECK(read_state(s));
    else if (get_while(s, &ex) == ICP_OK)
    {
        ICP_CUDA(cudaGraphLaunch(ex, s->ctx->stream));
        ICP_CHECK(read_state(s));
    }
    else
    {
        // fallback: one iteration per graph launch + blocking read of the state (what the reference does)
        ICP_CHECK(get_unrolled(s, 1, &ex));
        for (int32_t left = bound; left > 0; --left)
        {
            ICP_CUDA(cudaGraphLaunch(ex, s->ctx->stream));
            ICP_CHECK(read_state(s));
            if (s->h_state->done) break;
        }
    }
    if (k_out) *k_out = s->h_state->k;
    return ICP_OK;
}

extern "C" int icp_step_get_state(icp_step *s, icp_state *h_out)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_get_state: init() first"); return ICP_ERR_ARG; }
    ICP_CHECK(read_state(s));
    memcpy(h_out, s->h_state, sizeof(icp_state));
    return ICP_OK;
}

extern "C" int icp_step_get_pose_matrix(icp_step *s, float *h_T16)
{ ICP_ENTER_OBJ(s);
    ICP_CHECK(read_state(s));
    const icp_state *st = s->h_state;
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) h_T16[i * 4 + j] = st->s * st->R[i * 3 + j];
        h_T16[i * 4 + 3] = st->t[i];
    }
    h_T16[12] = h_T16[13] = h_T16[14] = 0.f; h_T16[15] = 1.f;
    return ICP_OK;
}

extern "C" void *icp_step_debug_ptr(icp_step *s, const char *name)
{
    if (!s->inited) return nullptr;
#define NAME(n, p) if (!strcmp(name, n)) return (void *)(p)
    NAME("reps", s->reps); NAME("rep_id", s->rep_id); NAME("N", s->N); NAME("O", s->O); NAME("perm", s->perm); NAME("Xp", s->Xp);
    NAME("Mt", s->Mt); NAME("q_rep", s->q_rep); NAME("qperm", s->qperm); NAME("Nq", s->Nq); NAME("Oq", s->Oq);
    NAME("Qp", s->Qp); NAME("NN", s->NN); NAME("NN_ID", s->NNID); NAME("W", s->W); NAME("sum_w", s->sum_w); NAME("mean", s->mean);
    NAME("DF", s->DF); NAME("DM", s->DM); NAME("S", s->S); NAME("Tk", s->Tk); NAME("state", s->state); NAME("F", s->F); NAME("M", s->M);
    NAME("T", s->T); NAME("evals", s->evals);
#undef NAME
    return fused_debug_ptr(s, name);
}

// one staged step with a CUDA event between the stages (the reference's run(timer), algorithms.hpp:2359-2399)
extern "C" int icp_step_run_timed(icp_step *s, float *h_ms7)
{ ICP_ENTER_OBJ(s);
    if (!s->inited) { icp_set_error("icp_step_run_timed: init() first"); return ICP_ERR_ARG; }
    ICP_CHECK(set_loop_params(s, 0, 0, 1, 0.0, 0.0));
    cudaEvent_t ev[8] = {};
    int rc = ICP_OK;
    for (int i = 0; i < 8 && rc == ICP_OK; ++i)
        if (cudaEventCreate(&ev[i]) != cudaSuccess) { ev[i] = nullptr; icp_set_error("icp_step_run_timed: cudaEventCreate failed"); rc = ICP_ERR_CUDA; }
    if (rc == ICP_OK) rc = staged_enqueue_iteration(s, s->ctx->stream, 0, 0, ev);
    if (rc == ICP_OK)
    {
        cudaEventSynchronize(ev[7]);
        for (int i = 0; i < 7; ++i) cudaEventElapsedTime(&h_ms7[i], ev[i], ev[i + 1]);
    }
    for (int i = 0; i < 8; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    return rc;
}

extern "C" int icp_step_stage1_executed(icp_step *s, uint64_t *e1x)
{ ICP_ENTER_OBJ(s);
    unsigned long long h = 0;
    ICP_CUDA(cudaMemcpyAsync(&h, s->evals + 2, sizeof(h), cudaMemcpyDeviceToHost, s->ctx->stream));
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    if (e1x) *e1x = h;
    return ICP_OK;
}

extern "C" int icp_step_stage2_executed(icp_step *s, uint64_t *e2x)
{ ICP_ENTER_OBJ(s);
    unsigned long long h = 0;
    ICP_CUDA(cudaMemcpyAsync(&h, s->evals + 3, sizeof(h), cudaMemcpyDeviceToHost, s->ctx->stream));
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    if (e2x) *e2x = h;
    return ICP_OK;
}

extern "C" int icp_step_eval_counts(icp_step *s, uint64_t *e1, uint64_t *e2)
{ ICP_ENTER_OBJ(s);
    unsigned long long h[2];
    ICP_CUDA(cudaMemcpyAsync(h, s->evals, sizeof(h), cudaMemcpyDeviceToHost, s->ctx->stream));
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    if (e1) *e1 = h[0];
    if (e2) *e2 = h[1];
    return ICP_OK;
}
