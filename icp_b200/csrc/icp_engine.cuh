// icp_engine.cuh -- registration engine object behind icp_step_* / icp_run (ICPStep<CR,CW>, ICP<CR,CW>).
#pragma once
#include "icp_stages.cuh"
#include <map>

// loop control block in device memory, written by the host before a (graph) launch
struct LoopParams
{
    double angle_thr, trans_thr;
    uint32_t check;            // 0: fixed iteration count, 1: ICP::check() thresholds
    uint32_t max_iterations;
    int32_t iters_left;        // fixed mode: remaining steps; check mode: safety bound
    uint32_t pad;
};

struct icp_step
{
    icp_ctx *ctx = nullptr;
    int rot_cfg = ICP_ROT_POWER_METHOD, w_cfg = ICP_W_WEIGHTED;
    int mode = ICP_MODE_FUSED;
    int rbc_mode = -1;          // mode whose buildRBC produced the current RBC (-1: none yet); the fused iteration kernels also
                                // need the acceleration tables only the FUSED build writes (nbr, wconst, nn_o, nnd, lb1)
    bool inited = false;
    uint32_t m = 0, nr = 0, lm_w = 0, lm_h = 0;
    float a = 0.f, c = 0.f, fg = 0.f, fp = 0.f;
    bool metric_override = false;
    bool count_evals = false;

    // bindable buffers (ICPStep::Memory)
    float *F = nullptr, *M = nullptr, *T = nullptr;
    bool own_F = false, own_M = false, own_T = false;

    // one arena for everything else
    void *arena = nullptr;
    size_t arena_bytes = 0;
    // RBC of the fixed set
    float *reps = nullptr, *Xp = nullptr;
    uint32_t *rep_id = nullptr, *N = nullptr, *O = nullptr, *perm = nullptr;
    // per iteration
    float *Mt = nullptr, *Qp = nullptr, *NN = nullptr, *W = nullptr, *mean = nullptr, *DF = nullptr, *DM = nullptr;
    float *S = nullptr, *Tk = nullptr, *Rk = nullptr;
    uint32_t *q_rep = nullptr, *qperm = nullptr, *Nq = nullptr, *Oq = nullptr;
    icp_dist_id *NNID = nullptr;
    double *sum_w = nullptr;
    DevState *state = nullptr;
    LoopParams *loop = nullptr;
    unsigned long long *evals = nullptr;   // [0] stage-1, [1] stage-2 evaluations (when count_evals)
    // scratch
    void *sort_scr = nullptr;
    float *red_f = nullptr; double *red_d = nullptr; float *mean_scr = nullptr; float *sij_part = nullptr, *sij_scr = nullptr;
    // fused-mode buffers
    void *fused = nullptr;

    // graphs
    std::map<uint32_t, cudaGraphExec_t> unrolled;      // key = n_iters (per mode, cleared on mode change)
    cudaGraphExec_t while_exec = nullptr;
    bool while_failed = false;
    cudaGraphExec_t build_exec = nullptr;
    LoopParams *h_loop = nullptr;                       // pinned staging for loop params
    icp_state *h_state = nullptr;                       // pinned staging for state read-back
};

int engine_enqueue_iteration(icp_step *s, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle);
int engine_enqueue_build(icp_step *s, cudaStream_t st);
void engine_drop_graphs(icp_step *s);

struct Carver
{
    char *base; size_t off;
    explicit Carver(void *b) : base((char *)b), off(0) {}
    template <typename Tp> Tp *take(size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        Tp *p = base ? (Tp *)(base + off) : nullptr;
        off += count * sizeof(Tp);
        return p;
    }
};

// fused mode (icp_fused.cu)
size_t fused_workspace_bytes(uint32_t m, uint32_t nr, int sm_count);
int fused_prepare(icp_step *s);
int fused_enqueue_build(icp_step *s, cudaStream_t st);
int fused_enqueue_iteration(icp_step *s, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle);
void *fused_debug_ptr(icp_step *s, const char *name);
int fused_invalidate(icp_step *s, cudaStream_t st, bool lane_order, bool bounds);
int fused_enqueue_persistent(icp_step *s, cudaStream_t st, uint32_t n_iters, int *ok);
