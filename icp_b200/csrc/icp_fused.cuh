// icp_fused.cuh -- the fused, batch-aware iteration kernels (the performance path).
//
// One ICP iteration of one frame pair = 4 launches in latency mode, 3 in batch mode (the reference needs >= 17 + a host round trip):
//   A  k_assign_tri   transform + nearest representative (exact triangle-inequality pruning; temporal pruning of the
//                     fallback points in the batch engine) + per-CTA stable ranks / histograms       [k_assign: exhaustive flavour]
//   B  k_colscan      exclusive prefix of the histograms over the CTAs (per representative), list sizes Nq   (latency mode)
//      k_colscan_sort the same in shared memory + list offsets Oq + the stable sorted order qperm            (batch mode)
//   C  k_search_grouped  queries grouped by representative inside the CTA, lists streamed through shared-memory tiles   (latency mode)
//      k_search_sorted   owns consecutive SORTED positions; exact temporal pruning of stage 2; with FUSE_D the last CTA of a
//                        pair continues with kernel D's body                                                              (batch mode)
//   D  k_reduce_solve / reduce_solve_body   sum(w) -> weighted means -> S matrix (reference tree shapes) -> rotation solve ->
//                     pose accumulation -> loop control            (latency mode: one 8-CTA thread-block cluster per pair;
//                     batch mode: one CTA per pair, its passes streaming through per-warp cp.async rings)
//      k_reduce_wide  the same body, one launch per pass over the whole GPU (one registration of > 10^5 points)
//   *  k_icp_persistent  phases A-D of one registration in ONE cooperative launch with software grid barriers (icp_run, long runs)
// 32-byte points move as 256-bit requests (LDG.E.ENL2.256).  One large image-ordered registration: kernel C works on 32 x 16
// patches of the landmark grid instead of runs of consecutive points (FusedCfg::ctile_w).
// Every kernel takes a table of per-pair pointers and uses blockIdx.y (A,B,C) / the cluster id (D) as the
// pair index, so the single-pair latency engine and the batched throughput engine share the same code.
#pragma once
#include "icp_engine.cuh"

struct PairPtrs
{
    const float *F;            // fixed landmarks   [m][8]
    const float *M;            // moving landmarks  [m][8]
    float *T;                  // D_IO_T {q,t,s}
    float *reps;               // [nr][8]
    float *Xp;                 // list-ordered fixed set [m][8]
    uint32_t *N, *O;           // list sizes / offsets of the fixed set
    uint32_t *rep_id, *perm;   // build outputs (per original point / list position -> original)
    uint32_t *q_rep;           // [m] representative of every query (original order)
    uint16_t *lrank;           // [m] stable rank of the query among equal reps inside its CTA chunk
    uint32_t *H;               // [nbA][nr] per-chunk histograms -> exclusive prefixes
    uint32_t *Nq, *Oq;         // [nr]
    uint32_t *wconst;          // [0] set by buildRBC: lanes 3 and 7 of every fixed point equal those of representative 0 (finite)
                               // [1] set by buildRBC: every representative-to-representative distance is finite (nbr is usable)
                               // [2] arrival counter of k_search_sorted<true>; [4..11] pose {q,t,s} of the previous iteration (kernel D)
                               // [12] tag of lperm (kernel A's seed-grouped lane order): iteration counter + 2 of the fused iteration whose
                               //      rank pass wrote it, i.e. valid exactly when it equals state->k + 1; 0 after buildRBC / a mode switch
                               // [13] the temporal-pruning bounds (nnd / nn_o, lb1 / tag1) describe the CURRENT moving set: set by
                               //      buildRBC (which resets them) and by kernel D, cleared by the single-pair engine at the start of
                               //      every run call (the caller may have rewritten M in between: the RBC only depends on F)
    uint32_t *nbx;             // [m][FUSED_NBX_K] per list position: its nearest points of the SAME list, ascending:
                               //   (distance chopped to bf16) << 16 | (position - list start); see k_list_neighbours
    uint32_t *nn_o;            // [m] per ORIGINAL query: list position of its nearest neighbour of the last iteration (the seed)
    float *nnd;                // [m] per original query: >= 0 NN distance found by kernel A's pruned walk; -1 = still to be searched
                               //     sorted flavour (settle): proven lower bound of sqrt(D) to every list point but nn_o; <= 0 = none
    uint2 *nbr;                // [nr][K] per representative: its K nearest other representatives {distance bits, index}, ascending
    uint32_t *qperm;           // [m] sorted position -> original query
    uint2 *QR;                 // [m] sorted flavour (Cmode 2): sorted position -> {original query, representative}, written by B'
    uint2 *nn2;                // [m] sorted flavour: per ORIGINAL query {bound (f32 bits, as nnd), list position of its last nearest neighbour
                               //     (as nn_o)} in ONE 8-byte word: one sector per gather instead of two
    float *Qs;                 // [m][8] span flavour (Cmode 3): the TRANSFORMED queries in sorted order, written by B'' (k_colscan_sort<.,true>)
    uint4 *Rs;                 // [m] span flavour: per sorted position {lower bound after this iteration's motion (f32 bits; <= 0: none),
                               //     list position of last iteration's nearest neighbour, original query index, representative}
    float *W;                  // [m]  weights, sorted order
    float *fxyz;               // [3][m] matched fixed points (NN.xyz), sorted order, SoA
    float *mxyz;               // [3][m] transformed queries (Q_p.xyz), sorted order, SoA
    icp_dist_id *NNID;         // [m] sorted order
    double *sum_w;
    float *mean;               // [8]
    float *S;                  // [11]
    float *Tk;                 // [8]
    float *Rk;                 // [9]
    DevState *state;
    LoopParams *loop;
    unsigned long long *evals; // [2] or NULL
    float *red;                // reduction scratch: see fused_red_elems()
    unsigned long long *prof;  // optional: clock64 stamps of kernel D's phases (rank 0), [7] = power iterations
    uint32_t *gbar;            // [4] arrival counter of the software grid barrier (persistent engine; zeroed before every launch)
};

struct FusedCfg
{
    uint32_t m, nr;
    uint32_t QB;        // queries per CTA chunk in kernel A (multiple of 32)
    uint32_t nbA;       // ceil(m / QB)
    int S;              // lanes per point group in kernel A (1..32)
    int QPT;            // points per group in kernel A (2 or 4)
    uint32_t TPB;       // threads per CTA of kernel A (512: two co-resident CTAs per SM, or 1024)
    int Amode;          // kernel A flavour: 0 = k_assign (every representative), 1 = k_assign_tri (triangle-inequality pruning)
    uint32_t K;         // neighbours kept per representative in PairPtrs::nbr (even, <= 32)
    uint32_t lm_w, lm_h;// landmark grid (single engine; 0 in the batch engine)
    uint32_t ctile_w;   // grouped kernel C: width of the CTA's patch of the landmark grid (0 = QG consecutive points)
    int nn_walk;        // kernel A also searches the nearest neighbour by a pruned walk from last iteration's match
    int SF;             // kernel A (pruned): lanes per point in the exhaustive pass of the unsettled points (8 or 32)
    int par_rank;       // kernel A ranks its chunk with all warps (needs ceil(QB/32)*nr*2 B of shared memory)
    int CL;             // cluster size of kernel D (1 or 8)
    int fastD;          // kernel D, CL = 8, m = 16384: points resident in shared memory, partials through distributed shared memory
    int TD;             // threads per CTA of kernel D (1024 with CL = 8; 256 / 512 / 1024 with CL = 1)
    int L;              // lanes per query in kernel C (1..32)
    uint32_t QC;        // queries per CTA in kernel C (multiple of 32)
    int Cmode;          // kernel C flavour: 0 = k_search<L> (L lanes per query, original order), 1 = k_search_grouped, 2 = k_colscan_sort + k_search_sorted,
                        //                   3 = k_colscan_sort (+ sorted query records) + k_search_span (lists staged in shared memory by bulk-async copies)
    uint32_t item_ovh;  // sorted flavour: fixed overhead of a work item in evaluations per lane (tail-decomposition rule; huge = never split)
    uint32_t span_pts;  // span flavour: capacity of the shared-memory list window of k_search_span, in points
    uint32_t QG;        // grouped C: consecutive queries per CTA (independent of kernel A's chunks)
    int aperm;          // kernel A: lane order of the pruned pass = the chunk's points grouped by last iteration's representative
    int pdl;            // latency mode: programmatic dependent launch along the kernel chain of an iteration
    int settle;         // sorted flavour, batch engine: exact temporal pruning of stage 2 (queries whose nearest neighbour provably did not change skip the list scan)
    int fuseD;          // sorted flavour, batch mode: kernel D runs in the tail of C' (last CTA of the pair), no separate launch
    int wideD;          // kernel D of one large registration as 3 whole-GPU launches + 1 single-CTA launch (k_reduce_wide) instead of a cluster
    int dring;          // kernel D, one CTA per pair (CL = 1, 512 threads, m % 2048 == 0): phases 2 / 3 stream through per-warp cp.async rings
    uint32_t GB;        // sorted flavour (Cmode 2): CTAs per pair of B' (k_colscan_sort)
    uint32_t TC;        // sorted flavour (Cmode 2): threads per CTA of C' (k_search_sorted), <= SORTED_WARPS * 32
    uint32_t QI;        // grouped C: queries per work item (8, 16 or 32); 32/QI lanes share one query's list
    float fg, fp, c;
    int weighted, power_method;
};

#define FUSED_NBR_K 32u
#define FUSED_NBX_K 16u
static inline size_t fused_nbr_elems(uint32_t nr) { return (size_t)nr * FUSED_NBR_K + 8; }   // uint2 elements

static inline size_t fused_red_elems(uint32_t m)
{
    // bs[nb128 + 4] | bm[6][nb128] | sp[11][nb512] ; + slack
    const size_t nb128 = div_up(m, 128), nb512 = div_up(div_up(m, 4), 512);
    return (nb128 + 8) + 6 * nb128 + 11 * (nb512 + 4) + 64;
}

void fused_choose_cfg(FusedCfg *cfg, uint32_t m, uint32_t nr, int sm_count, uint32_t n_pairs);
int fused_launch_build(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, uint32_t lm_w, uint32_t lm_h);
int fused_launch_iteration(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs,
                           cudaGraphConditionalHandle handle, int use_handle);
