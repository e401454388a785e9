// icp_stages.cuh -- internal launch functions of the per-stage kernels (icp_stages.cu).
// All take explicit scratch pointers and a stream so that the engine can capture them in a CUDA graph.
#pragma once
#include "icp_common.cuh"

// scratch sizing helpers ---------------------------------------------------------------------------
struct SortScratch   // stable counting sort of n keys in [0,nr)
{
    uint32_t *H;        // [nb][nr] per-CTA histograms -> exclusive prefix over CTAs
    uint16_t *lrank;    // [n] rank of the element among equal keys inside its CTA
    static size_t bytes(uint32_t n, uint32_t nr) { return (size_t)div_up(n, 256) * nr * 4 + (size_t)((n + 1) & ~1u) * 2 + 64; }
    void carve(void *base, uint32_t n, uint32_t nr)
    {
        H = (uint32_t *)base;
        lrank = (uint16_t *)((char *)base + (size_t)div_up(n, 256) * nr * 4);
    }
};

static inline size_t reduce_scratch_elems(uint32_t n) { return (size_t)div_up(n, 128) + 8; }

int launch_get_lms(cudaStream_t st, const float *cloud, float *lms);
int launch_rgbd_to_pc8d(cudaStream_t st, int sm_count, const uint16_t *depth, const uint8_t *rgb, uint32_t W, uint32_t H, float focal, float *cloud);
int launch_get_reps(cudaStream_t st, const float *lms, uint32_t W, uint32_t H, uint32_t nr, float *reps);
int launch_transform_q(cudaStream_t st, const float *M, const float *T8, float *out, uint32_t m);
int launch_transform_m(cudaStream_t st, const float *M, const float *T16, float *out, uint32_t m);
// nearest representative; T8 != NULL => fused ICPTransform<QUATERNION> (transformed points written to Xt_out if non-NULL)
int launch_nearest_rep(cudaStream_t st, int sm_count, const float *X, uint32_t n, const float *R, uint32_t nr,
                       float fg, float fp, const float *T8, float *Xt_out, uint32_t *rep_id);
// stable counting sort: N, O, perm (pos -> original index); Xsrc/Xp optional gather of the 8-D points
int launch_counting_sort(cudaStream_t st, const uint32_t *key, uint32_t n, uint32_t nr, SortScratch sc,
                         uint32_t *N, uint32_t *O, uint32_t *perm, const float *Xsrc, float *Xp);
int launch_rbc_stage2(cudaStream_t st, const float *Q, const uint32_t *q_rep, const uint32_t *qperm, uint32_t m,
                      const float *Xp, uint32_t n_db, const uint32_t *O, const uint32_t *N, float fg, float fp,
                      float *Qp, float *NN, icp_dist_id *NNID, unsigned long long *evals);
// scratchF: >= reduce_scratch_elems(n)+4 floats; scratchD: 2 x reduce_scratch_elems(n) doubles
int launch_weights(cudaStream_t st, const icp_dist_id *in, float *W, double *sum_w, uint32_t n, float *scratchF, double *scratchD);
// scratch: 2 x (2*4*reduce_scratch_elems(n)) floats
int launch_mean(cudaStream_t st, const float *F, const float *M, const float *W, const double *sum_w, float *mean,
                uint32_t n, float *scratch);
int launch_devs(cudaStream_t st, const float *F, const float *M, const float *mean, float *DF, float *DM, uint32_t n);
// partials: 11*G floats (G = ceil(m/4)); scratch: 2 x 11 x (G/512+8) floats
int launch_sij(cudaStream_t st, const float *DM, const float *DF, const float *W, float *S11, uint32_t m, float c,
               float *partials, float *scratch);
static inline size_t sij_partials_elems(uint32_t m) { return (size_t)11 * div_up(m, 4); }
static inline size_t sij_scratch_elems(uint32_t m) { return (size_t)2 * 11 * (div_up(div_up(m, 4), 512) + 8); }
int launch_power_method(cudaStream_t st, const float *S11, const float *mean, float *Tk8);
int launch_svd_solve(cudaStream_t st, const float *S11, const float *mean, float *Tk8, float *Rk9);
// pose accumulation (algorithms.cpp:4683-4697): state := f(state, Tk[, Rk]); T8 := {q,t,s}; k++;
// done := converged (ICP::check thresholds) when check != 0.
int launch_accumulate(cudaStream_t st, DevState *state, const float *Tk8, const float *Rk9, float *T8,
                      int check, uint32_t max_iterations, double angle_thr, double trans_thr);
int launch_reduce_sum_f(cudaStream_t st, const float *in, uint32_t cols, uint32_t rows, float *out, float *scratch);
static inline size_t reduce_sum_scratch_elems(uint32_t cols, uint32_t rows) { return (size_t)2 * rows * (div_up(cols, 512) + 8); }
int launch_reduce_min_f(cudaStream_t st, const float *in, uint32_t cols, uint32_t rows, float *out);
int launch_reduce_max_ui(cudaStream_t st, const uint32_t *in, uint32_t cols, uint32_t rows, uint32_t *out);
int launch_scan_i(cudaStream_t st, const int32_t *in, uint32_t cols, uint32_t rows, int inclusive, int32_t *out);

#ifdef __CUDACC__
// device routines shared with the fused engine ------------------------------------------------------
// Single-thread solvers (icp_solve.cuh)
#endif
