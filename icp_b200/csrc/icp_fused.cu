// icp_fused.cu -- fused, batch-aware iteration kernels (see icp_fused.cuh for the pipeline overview).
// Results are bit-identical to the staged kernels (icp_stages.cu) and to oracle/icp_oracle.cpp: same
// per-element arithmetic, same stable rep-sorted order, same reduction tree shapes.
#include "icp_fused.cuh"
#include "icp_solve.cuh"
#include <cooperative_groups.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
namespace cg = cooperative_groups;

// 32-byte points move as ONE 256-bit request (icp_common.cuh); ICP_FUSED_LD256=0 builds the 2 x 128-bit flavour for the A/B
#ifndef ICP_FUSED_LD256
#define ICP_FUSED_LD256 1
#endif
#if ICP_FUSED_LD256
#define ld_pt8 ld_pt8_v8
#define ld_pt8_cg ld_pt8_cg_v8
#define st_pt8 st_pt8_v8
#endif
#ifdef ICP_T_LDG
#define ICP_LOAD_T __ldg
#else
#define ICP_LOAD_T __ldcg
#endif
#define TPB_A 1024           // upper bound of kernel A's block size (the actual size is cfg.TPB)
#define TPB_D 1024
#ifndef ASSIGN_MINB4
#define ASSIGN_MINB4 2      // resident 512-thread CTAs per SM the QPT = 4 flavour of kernel A is compiled for (2 => 64 registers)
#endif

// =================================================================================================
// A: nearest representative (+ fused ICPTransform<QUATERNION> when SEARCH) and stable in-chunk ranks.
// CTA = one chunk of QB consecutive points; S adjacent lanes share a point and scan nr/S representatives
// each out of shared memory (broadcast LDS.128); ordered argmin merge by warp shuffle.
// =================================================================================================
// Every CTA first checks that the two homogeneous lanes (w of xyz1 and of rgb1) are the same finite constants in
// all representatives, and every warp that they are the same in its points: then dw = da = +0 exactly and the two
// terms (dw^2, da^2) add +0 to non-negative partial sums -- dropping them is bit-exact (dist6 below).  pc8d clouds
// always satisfy this (lanes 3 and 7 are 1); anything else takes the full 8-lane path.
__device__ __forceinline__ float dist6(const float4 &qlo, const float4 &qhi, const float4 &xlo, const float4 &xhi, float fg, float fp)
{
    float d0 = __fsub_rn(qlo.x, xlo.x), d1 = __fsub_rn(qlo.y, xlo.y), d2 = __fsub_rn(qlo.z, xlo.z);
    float d4 = __fsub_rn(qhi.x, xhi.x), d5 = __fsub_rn(qhi.y, xhi.y), d6 = __fsub_rn(qhi.z, xhi.z);
    float g = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
    float p = __fadd_rn(__fadd_rn(__fmul_rn(d4, d4), __fmul_rn(d5, d5)), __fmul_rn(d6, d6));
    return __fadd_rn(__fmul_rn(fg, g), __fmul_rn(fp, p));
}
__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) < CUDART_INF_F; }

// Scan of the representatives c, c+S, ... by one lane for its QPT points, with two exact shortcuts:
//  * seed: (best, bi) starts from the distance to a guessed representative (bi[] on entry) instead of +inf.  The update
//    rule is the ORDERED compare (smaller distance, then smaller index), so the final (min, lowest index) pair is the
//    one a sequential strict-'<' scan from +inf finds, whatever the seed was.
//  * partial-distance early-out: d = fl (pg + fl (fp * p)) with pg = fl (fg * g) and fl (fp * p) >= +0 (fp >= 0), and
//    rounding is monotonic, so d >= pg.  If pg > best for every point of every lane of the warp, no lane can update and
//    the colour half of the metric (10 of its 19 operations) is skipped.  A NaN pg compares false: its d is NaN too and
//    would not have been selected either.
// FAST = the two homogeneous lanes are constant (see dist6).
template <int S, int QPT, bool FAST>
__device__ __forceinline__ void scan_reps(const float4 *__restrict__ sRlo, const float4 *__restrict__ sRhi, uint32_t nr, uint32_t c,
                                          const pt8 (&q)[QPT], float (&best)[QPT], uint32_t (&bi)[QPT], float fg, float fp, bool prune)
{
#pragma unroll
    for (int j = 0; j < QPT; ++j)
    {
        const float4 rlo = sRlo[bi[j]], rhi = sRhi[bi[j]];
        const float d = FAST ? dist6(q[j].lo, q[j].hi, rlo, rhi, fg, fp) : dist8(q[j].lo, q[j].hi, rlo, rhi, fg, fp);
        if (d < CUDART_INF_F && prune) best[j] = d;
        else { best[j] = CUDART_INF_F; bi[j] = c; }
    }
#pragma unroll 2
    for (uint32_t r = c; r < nr; r += S)
    {
        const float4 rlo = sRlo[r];
        float pg[QPT];
        bool pass = false;
#pragma unroll
        for (int j = 0; j < QPT; ++j)
        {
            const float d0 = __fsub_rn(q[j].lo.x, rlo.x), d1 = __fsub_rn(q[j].lo.y, rlo.y), d2 = __fsub_rn(q[j].lo.z, rlo.z);
            float g = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            if (!FAST) { const float d3 = __fsub_rn(q[j].lo.w, rlo.w); g = __fadd_rn(g, __fmul_rn(d3, d3)); }
            pg[j] = __fmul_rn(fg, g);
            pass = pass || (pg[j] <= best[j]);
        }
        if (__any_sync(FULL_MASK, pass || !prune))
        {
            const float4 rhi = sRhi[r];
#pragma unroll
            for (int j = 0; j < QPT; ++j)
            {
                const float d4 = __fsub_rn(q[j].hi.x, rhi.x), d5 = __fsub_rn(q[j].hi.y, rhi.y), d6 = __fsub_rn(q[j].hi.z, rhi.z);
                float p = __fadd_rn(__fadd_rn(__fmul_rn(d4, d4), __fmul_rn(d5, d5)), __fmul_rn(d6, d6));
                if (!FAST) { const float d7 = __fsub_rn(q[j].hi.w, rhi.w); p = __fadd_rn(p, __fmul_rn(d7, d7)); }
                const float d = __fadd_rn(pg[j], __fmul_rn(fp, p));
                if (d < best[j] || (d == best[j] && r < bi[j])) { best[j] = d; bi[j] = r; }
            }
        }
    }
}

// Latency-mode flavour of scan_reps (S lanes per point): the partial-distance early-out makes every trip wait for a warp
// vote on the previous trip's result (8 dependent trips of ~145 cycles at 256 representatives); here 8 evaluations are
// independent of each other (one LDS latency + one 19-operation chain for all of them) and only the ordered selects are
// sequential.  Same candidates, same ordered update rule => the same (minimum, lowest index) pair.
template <int S, bool FAST>
__device__ __forceinline__ void scan_reps_ilp(const float4 *__restrict__ sRlo, const float4 *__restrict__ sRhi, uint32_t nr, uint32_t c,
                                                const pt8 &q, float &best, uint32_t &bi, float fg, float fp, bool prune)
{
    {
        const float4 rlo = sRlo[bi], rhi = sRhi[bi];
        const float d = FAST ? dist6(q.lo, q.hi, rlo, rhi, fg, fp) : dist8(q.lo, q.hi, rlo, rhi, fg, fp);
        if (d < CUDART_INF_F && prune) best = d;
        else { best = CUDART_INF_F; bi = c; }
    }
    for (uint32_t r0 = c; r0 < nr; r0 += 8u * S)
    {
        float d[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u)
        {
            const uint32_t r = min(r0 + (uint32_t)S * u, nr - 1u);
            const float4 rlo = sRlo[r], rhi = sRhi[r];
            d[u] = FAST ? dist6(q.lo, q.hi, rlo, rhi, fg, fp) : dist8(q.lo, q.hi, rlo, rhi, fg, fp);
        }
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u)
        {
            const uint32_t r = r0 + (uint32_t)S * u;
            if (r < nr && (d[u] < best || (d[u] == best && r < bi))) { best = d[u]; bi = r; }
        }
    }
}

// Stable ranks of the chunk's points among equal representatives + per-chunk histogram (tail of kernel A).
// keys[l] = representative of local point l.  Whole CTA; starts with a barrier.
__device__ __forceinline__ uint32_t cta_exscan_smem(const uint32_t *in_s, uint32_t n, uint32_t *out_s, uint32_t *warp_tot);

// lperm_out (optional, parallel path only): the chunk's points ordered by representative (stable), as local indices; the next
// iteration's pruned pass hands them to the lanes in this order, so that a warp walks the neighbourhood of ONE seed.
// scratch: >= nr u32 of shared memory (only used with lperm_out).
__device__ __forceinline__ void chunk_rank_store(const PairPtrs &P, const FusedCfg &cfg, uint32_t *keys, uint32_t *cnt, uint16_t *slc,
                                                 uint32_t *q_rep, uint32_t q0, uint32_t nq, const uint32_t bx /* chunk = (virtual) block index */,
                                                 uint16_t *lperm_out = nullptr, uint32_t *scratch = nullptr)
{
    const uint32_t nr = cfg.nr, QB = cfg.QB, TPB = blockDim.x, tid = threadIdx.x;
    const uint32_t nsl = (QB + 31u) / 32u;
    const bool par_rank = cfg.par_rank != 0;
    __syncthreads();
    if (par_rank)
    {
        // stable ranks inside the chunk, all warps: rank inside the 32-point slice by match_any, per-slice counts,
        // then an exclusive prefix over the slices per representative
        const uint32_t lane = tid & 31u, nw = TPB >> 5;
        for (uint32_t sl = tid >> 5; sl < nsl; sl += nw)
        {
            const uint32_t l = sl * 32u + lane;
            const bool v = l < nq;
            const uint32_t k = v ? keys[l] : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(FULL_MASK, k);
            const uint32_t lr = __popc(peers & lanemask_lt());
            if (v)
            {
                keys[l] = k | (lr << 16);
                if (lr == 0) slc[sl * nr + k] = (uint16_t)__popc(peers);
            }
        }
        __syncthreads();
        for (uint32_t r = tid; r < nr; r += TPB)
        {
            uint32_t run = 0;
            for (uint32_t sl = 0; sl < nsl; ++sl) { const uint32_t t = slc[sl * nr + r]; slc[sl * nr + r] = (uint16_t)run; run += t; }
            P.H[(size_t)bx * nr + r] = run;
            if (lperm_out) cnt[r] = run;
        }
        __syncthreads();
        if (lperm_out)
        {
            __shared__ uint32_t rk_tot[32];
            cta_exscan_smem(cnt, nr, scratch, rk_tot);       // first position of every representative inside the chunk
        }
        for (uint32_t l = tid; l < nq; l += TPB)
        {
            const uint32_t kk = keys[l], k = kk & 0xFFFFu;
            const uint32_t lr = slc[(l >> 5) * nr + k] + (kk >> 16);
            P.lrank[q0 + l] = (uint16_t)lr;
            q_rep[q0 + l] = k;
            if (lperm_out) lperm_out[q0 + scratch[k] + lr] = (uint16_t)l;
        }
        // lperm is valid for the NEXT iteration only: the tag is the iteration counter that iteration will see + 1.  (A plain
        // flag set here would also be seen by CTAs of THIS launch that start later -- a second wave -- before their chunk is written.)
        if (lperm_out && bx == 0 && tid == 0) P.wconst[12] = __ldcg(&P.state->k) + 2u;
        return;
    }
    // serial fallback (shared memory too small for the per-slice counts): warp 0 walks the chunk 32 points at a time
    if (tid < 32)
    {
        for (uint32_t g0 = 0; g0 < nq; g0 += 32)
        {
            const uint32_t l = g0 + tid;
            const bool v = l < nq;
            const uint32_t k = v ? keys[l] : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(FULL_MASK, k);
            const uint32_t lr = __popc(peers & lanemask_lt());
            uint32_t base = 0;
            if (v)
            {
                base = cnt[k];
                P.lrank[q0 + l] = (uint16_t)(base + lr);
                q_rep[q0 + l] = k;
            }
            __syncwarp();
            if (v && lr == 0) cnt[k] = base + __popc(peers);
            __syncwarp();
        }
    }
    __syncthreads();
    for (uint32_t r = tid; r < nr; r += TPB) P.H[(size_t)bx * nr + r] = cnt[r];
}

template <int S, int QPT, bool SEARCH>
__global__ void __launch_bounds__(QPT == 4 ? 512 : TPB_A, QPT == 4 ? ASSIGN_MINB4 : 1) k_assign(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ float4 smem_a[];
    const uint32_t nr = cfg.nr, m = cfg.m, QB = cfg.QB, TPB = blockDim.x;
    float4 *sRlo = smem_a;                                       // [nr] xyz1 halves
    float4 *sRhi = sRlo + nr;                                    // [nr] rgb1 halves
    uint32_t *keys = reinterpret_cast<uint32_t *>(sRhi + nr);    // [QB]  representative | (rank inside the 32-point slice << 16)
    uint32_t *cnt = keys + QB;                                   // [nr]
    uint16_t *slc = reinterpret_cast<uint16_t *>(cnt + nr);      // [ceil(QB/32)][nr] per-slice counts (parallel ranking only)
    const PairPtrs P = table[blockIdx.y];
    if (SEARCH && P.state->done) return;
    const uint32_t tid = threadIdx.x;
    const uint32_t nsl = (QB + 31u) / 32u;
    const bool par_rank = cfg.par_rank != 0;
    const float4 r0lo = __ldg((const float4 *)P.reps), r0hi = __ldg((const float4 *)P.reps + 1);
    bool okw = finite_f(r0lo.w) && finite_f(r0hi.w);
    for (uint32_t i = tid; i < nr * 2u; i += TPB)
    {
        const float4 v = __ldg((const float4 *)P.reps + i);
        if (i & 1u) { sRhi[i >> 1] = v; okw = okw && (v.w == r0hi.w); }
        else { sRlo[i >> 1] = v; okw = okw && (v.w == r0lo.w); }
    }
    for (uint32_t i = tid; i < nr; i += TPB) cnt[i] = 0u;
    if (par_rank) for (uint32_t i = tid; i < (nsl * nr + 1u) / 2u; i += TPB) reinterpret_cast<uint32_t *>(slc)[i] = 0u;
    const bool reps_w_const = __syncthreads_and(okw) != 0;

    const float *X = SEARCH ? P.M : P.F;
    const uint32_t q0 = blockIdx.x * QB;
    const uint32_t nq = min(QB, m - q0);
    float4 tq, tt;
    if (SEARCH) { tq = __ldg((const float4 *)P.T); tt = __ldg((const float4 *)P.T + 1); }
    const float fg = cfg.fg, fp = cfg.fp;
    const bool prune = fp >= 0.f;
    // stage-1 temporal pruning of the points that need the exhaustive scan (batch engine, metric weights in [0, 1])
    const bool settle1 = SEARCH && cfg.settle != 0 && cfg.nn_walk == 0 && fg >= 0.f && fg <= 1.f && fp >= 0.f && fp <= 1.f;
    uint32_t k_now = 0u;
    uint32_t *q_rep = SEARCH ? P.q_rep : P.rep_id;
    // S adjacent lanes form a group that owns QPT consecutive points; lane c of the group scans the
    // representatives c, c+S, c+2S, ... (the S lanes read S consecutive 16-byte halves: conflict-free
    // LDS.128), and every representative fetched from shared memory is reused for the QPT points held in
    // registers (divides the LDS traffic per distance evaluation by QPT).
    const uint32_t GROUPS = TPB / S;
    for (uint32_t t0 = 0; t0 < nq; t0 += GROUPS * QPT)
    {
        const uint32_t ql0 = t0 + (tid / S) * QPT, c = tid % S;
        pt8 q[QPT];
        float best[QPT];
        uint32_t bi[QPT];
        bool fast = reps_w_const;
#pragma unroll
        for (int j = 0; j < QPT; ++j)
        {
            const bool valid = ql0 + j < nq;
            const uint32_t gi = valid ? q0 + ql0 + j : q0;
            q[j] = ld_pt8(X, gi);
            if (SEARCH) q[j].lo = transform_q_xyz(q[j].lo, tq, tt);
            fast = fast && (q[j].lo.w == r0lo.w) && (q[j].hi.w == r0hi.w);
            // seed: the representative this point had last time (any valid index works: the result never depends on it)
            bi[j] = min(__ldcg(q_rep + gi), nr - 1u);
        }
        const bool warp_fast = __all_sync(FULL_MASK, fast);
        if (!SEARCH && !warp_fast && (tid & 31u) == 0) *P.wconst = 0u;
        if (warp_fast) scan_reps<S, QPT, true>(sRlo, sRhi, nr, c, q, best, bi, fg, fp, prune);
        else scan_reps<S, QPT, false>(sRlo, sRhi, nr, c, q, best, bi, fg, fp, prune);
#pragma unroll
        for (int j = 0; j < QPT; ++j)
        {
            float b = best[j];
            uint32_t id = bi[j];
#pragma unroll
            for (int off = 1; off < S; off <<= 1)
            {
                const float od = __shfl_xor_sync(FULL_MASK, b, off);
                const uint32_t oi = __shfl_xor_sync(FULL_MASK, id, off);
                if (od < b || (od == b && oi < id)) { b = od; id = oi; }
            }
            if (c == 0 && ql0 + j < nq) keys[ql0 + j] = (b < CUDART_INF_F) ? id : 0u;
        }
    }
    chunk_rank_store(P, cfg, keys, cnt, slc, q_rep, q0, nq, blockIdx.x);
}

// =================================================================================================
// Representative neighbour table (buildRBC): for every representative s, its FUSED_NBR_K nearest other
// representatives, ascending by the RBC metric (ties: lower index first), as {distance bits, index}.
// One CTA per representative; ranks by counting (nr <= 4096).
// =================================================================================================
__global__ void __launch_bounds__(256) k_rep_neighbours(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ float sD[];                    // [nr]
    const PairPtrs P = table[blockIdx.y];
    const uint32_t nr = cfg.nr, K = cfg.K, s = blockIdx.x;
    const pt8 rs = ld_pt8(P.reps, s);
    bool bad = false;
    for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x)
    {
        const pt8 x = ld_pt8(P.reps, r);
        const float d = dist8(rs.lo, rs.hi, x.lo, x.hi, cfg.fg, cfg.fp);
        bad = bad || !(d < CUDART_INF_F);            // NaN or overflow: the triangle bound cannot be trusted
        sD[r] = (r == s) ? -1.f : d;                 // s itself sorts first and is not stored
    }
    if (bad) P.wconst[1] = 0u;
    __syncthreads();
    for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x)
    {
        if (r == s) continue;
        const float d = sD[r];
        uint32_t rank = 0;
        for (uint32_t r2 = 0; r2 < nr; ++r2)
        {
            const float d2 = sD[r2];
            rank += (d2 < d || (d2 == d && r2 < r)) ? 1u : 0u;
        }
        if (rank >= 1u && rank - 1u < K) P.nbr[(size_t)s * K + (rank - 1u)] = make_uint2(__float_as_uint(d), r);
    }
}

// =================================================================================================
// A (pruned): nearest representative by triangle-inequality pruning -- same result as k_assign, a fraction of the
// distance evaluations.
//   The RBC metric is a squared Euclidean distance in a scaled space (fg, fp >= 0), so for a point p, a guessed
//   representative s (its representative of the previous iteration) and any other representative r:
//        sqrt D(p,r) >= sqrt D(s,r) - sqrt D(p,s).
//   If sqrt D(s,r) > sqrt D(p,s) + sqrt best then D(p,r) > best: r cannot be the nearest, nor tie with it (best = the
//   smallest distance found so far, D(p,s) at the start).  (a + b)^2 <= 2 (a^2 + b^2), and with floating-point distances
//   (relative error <= 9 ulp, absolute error < 1e-36 from underflow, fg, fp in [0,1]) the test used is
//        D~(s,r) > 2 (D~(p,s) + best) (1 + 1e-3) + 1e-30,
//   which implies D~(p,r) > best strictly (DESIGN.md section 4 has the error analysis).  The lane walks the sorted
//   neighbour row of s and stops at the first entry that fails the test (best only shrinks => later entries fail too).
//   Every candidate that is evaluated is evaluated with the exact reference arithmetic and compared with the
//   ordered rule (smaller distance, then smaller index), so the winner is the one the full strict-'<' scan finds.
//   Points whose bound cannot be proven inside the K stored neighbours (outliers, bad guesses, non-finite data)
//   are collected and scanned against every representative by 8 lanes each (scan_reps: seeded + early-out).
// CTA = one chunk of QB points, one point per lane in the pruned pass.
// =================================================================================================
// optional timeline stamps (latency-mode diagnosis; P.prof is NULL in the batch engine): slot 16 + 8*kernel + i
__device__ __forceinline__ unsigned long long gtime_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// Programmatic dependent launch (latency mode): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor still runs; pdl_wait() blocks until the predecessor has completed and its writes are visible,
// pdl_trigger() lets the NEXT kernel of the stream begin its launch.  Both are no-ops for ordinary launches.  Every fused
// iteration kernel does wait-then-trigger before anything else, so only launch latency is overlapped, never data.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// software prefetch into L2 / L1 (no destination register): used where a thread works through several gathers one after the
// other -- the addresses of all of them are issued first, the dependent loads then hit the cache
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
#define PROF_STAMP(P, kid, i, val) do { if ((P).prof && blockIdx.x == 0 && threadIdx.x == 0) (P).prof[16 + 8 * (kid) + (i)] = (val); } while (0)
// slot 48 + kernel: latest end of ANY CTA of the kernel (all iterations so far => the last iteration's last CTA)
#define PROF_END_ALL(P, kid) do { if ((P).prof && threadIdx.x == 0) atomicMax((P).prof + 48 + (kid), gtime_ns()); } while (0)
#define TRI_S 8
// exclusion threshold on D~(s,r): 2 (D(p,s) + best) >= (sqrt D(p,s) + sqrt best)^2 (equal when best == D(p,s), the usual case),
// inflated by the rounding slack (see the header above).  No square root: the kernel stays free of FFMA sequences.
__device__ __forceinline__ float tri_thr(float ds, float best)
{
    return __fadd_rn(__fmul_rn(__fmul_rn(2.f, __fadd_rn(ds, best)), 1.001f), 1e-30f);
}

// walk of the sorted neighbour row of the guessed representative s (see k_assign_tri); returns the nearest representative
// SEC (temporal pruning of stage 1): the walk also leaves what is needed for a lower bound of sqrt(D) between the
// point and EVERY representative but the winner -- sec: the second smallest evaluated distance (the seed counts as
// evaluated); dstop: D~(s, e) of the row entry e at which the walk stopped.  Every representative that was not evaluated lies
// at D~(s, r) >= dstop from the seed (the row is sorted, representatives beyond the row are farther still), hence at
// sqrt D(p, r) >= sqrt D(s, r) - sqrt D(p, s) from the point.
template <bool FAST, bool SEC>
__device__ __forceinline__ uint32_t tri_walk(const uint2 *__restrict__ row, uint32_t K, const pt8 &q, const float4 *sRlo, const float4 *sRhi,
                                             float fg, float fp, float ds, uint32_t s, uint32_t &ecnt, float &sec, float &dstop)
{
    float best = ds, thr = tri_thr(ds, ds);
    uint32_t bi = s;
    // the last entry decides whether the walk is guaranteed to stop inside the stored row (else: 0xFFFFFFFF = scan everything);
    // it is fetched together with the first batch (one memory latency)
    const float last = __uint_as_float(__ldg(&row[K - 1u].x));
    // the row is fetched 8 entries (4 x 16 bytes, one memory latency) at a time; most walks end inside the first batch
    for (uint32_t k0 = 0; k0 < K; k0 += 8u)
    {
        uint4 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            e[j] = (k0 + 2u * j < K) ? __ldg(reinterpret_cast<const uint4 *>(row + k0) + j) : make_uint4(0x7f800000u, 0u, 0x7f800000u, 0u);
        if (k0 == 0u && !(last > thr)) return 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            if (__uint_as_float(e[j].x) > thr) { if (SEC) dstop = __uint_as_float(e[j].x); return bi; }
            {
                const uint32_t r = e[j].y;
                ++ecnt;
                const float d = FAST ? dist6(q.lo, q.hi, sRlo[r], sRhi[r], fg, fp) : dist8(q.lo, q.hi, sRlo[r], sRhi[r], fg, fp);
                if (SEC) sec = fminf(sec, fmaxf(d, best));
                if (d < best || (d == best && r < bi)) { best = d; bi = r; thr = tri_thr(ds, best); }
            }
            if (__uint_as_float(e[j].z) > thr) { if (SEC) dstop = __uint_as_float(e[j].z); return bi; }
            {
                const uint32_t r = e[j].w;
                ++ecnt;
                const float d = FAST ? dist6(q.lo, q.hi, sRlo[r], sRhi[r], fg, fp) : dist8(q.lo, q.hi, sRlo[r], sRhi[r], fg, fp);
                if (SEC) sec = fminf(sec, fmaxf(d, best));
                if (d < best || (d == best && r < bi)) { best = d; bi = r; thr = tri_thr(ds, best); }
            }
        }
    }
    if (SEC) dstop = -1.f;           // not reached (the last entry exceeds thr and thr never grows): no bound
    return bi;
}

// =================================================================================================
// List neighbour table (buildRBC): for every list position k, the FUSED_NBX_K nearest points of the SAME list, ascending by
// the RBC metric (ties: lower position first), packed as (distance chopped to bf16) << 16 | (position - list start).
// Chopping rounds the distance DOWN, which only makes the exclusion test of nn_walk more conservative; the chopped values
// stay sorted.  Padding = +inf; a row that cannot be trusted (NaN distance, list longer than 65535) gets -inf in its last
// entry, which sends every query seeded there to the exhaustive scan.  One thread per list position.
// =================================================================================================
__global__ void __launch_bounds__(128) k_list_neighbours(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    const PairPtrs P = table[blockIdx.y];
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cfg.m) return;
    const uint32_t r = __ldcg(P.rep_id + __ldcg(P.perm + k));
    const uint32_t o = __ldcg(P.O + r), n = __ldcg(P.N + r);
    const pt8 x = ld_pt8_cg(P.Xp, k);
    float td[FUSED_NBX_K];
    uint32_t ti[FUSED_NBX_K];
#pragma unroll
    for (int t = 0; t < (int)FUSED_NBX_K; ++t) { td[t] = CUDART_INF_F; ti[t] = 0xFFFFu; }
    bool bad = n > 65535u;
    if (!bad)
        for (uint32_t j = o; j < o + n; ++j)
        {
            if (j == k) continue;
            const pt8 y = ld_pt8_cg(P.Xp, j);
            const float d = dist8(x.lo, x.hi, y.lo, y.hi, cfg.fg, cfg.fp);
            bad = bad || (d != d);
            if (d < td[FUSED_NBX_K - 1])
            {
                td[FUSED_NBX_K - 1] = d; ti[FUSED_NBX_K - 1] = j - o;
#pragma unroll
                for (int t = (int)FUSED_NBX_K - 1; t > 0; --t)
                    if (td[t] < td[t - 1])          // strict: an equal distance met later stays behind (lower position first)
                    {
                        const float fd = td[t]; td[t] = td[t - 1]; td[t - 1] = fd;
                        const uint32_t fi = ti[t]; ti[t] = ti[t - 1]; ti[t - 1] = fi;
                    }
            }
        }
    uint32_t e[FUSED_NBX_K];
#pragma unroll
    for (int t = 0; t < (int)FUSED_NBX_K; ++t) e[t] = (__float_as_uint(td[t]) & 0xFFFF0000u) | ti[t];
    if (bad) e[FUSED_NBX_K - 1] = 0xFF80FFFFu;
    uint4 *row = reinterpret_cast<uint4 *>(P.nbx + (size_t)k * FUSED_NBX_K);
#pragma unroll
    for (int t = 0; t < (int)FUSED_NBX_K / 4; ++t) row[t] = make_uint4(e[4 * t], e[4 * t + 1], e[4 * t + 2], e[4 * t + 3]);
}

// Nearest neighbour of a query inside its representative's list by the same triangle bound as tri_walk, anchored at the
// query's match of the previous iteration (nn_o): evaluate the anchor, then only the anchor's list neighbours x with
// D~(anchor, x) <= 2 (D(q, anchor) + best) (1 + 1e-3) + 1e-30; everything else is provably farther than best.  Ordered
// compare (smaller distance, then lower list position) == the sequential strict-'<' scan of the list.  Returns the
// distance (and stores the position) when the walk is conclusive, -1 when the exhaustive scan of kernel C has to decide:
// anchor outside the current list, non-finite anchor distance, bound not provable inside the stored row.
template <bool FAST>
__device__ __forceinline__ float nn_walk(const PairPtrs &P, const pt8 &q, uint32_t r, uint32_t gi, float fg, float fp, uint32_t &ecnt)
{
    const uint32_t o = __ldg(P.O + r), n = __ldg(P.N + r);
    const uint32_t sp = __ldcg(P.nn_o + gi);
    if (sp - o >= n) return -1.f;
    const pt8 xs = ld_pt8(P.Xp, sp);
    const float ds = FAST ? dist6(q.lo, q.hi, xs.lo, xs.hi, fg, fp) : dist8(q.lo, q.hi, xs.lo, xs.hi, fg, fp);
    ++ecnt;
    if (!(ds < CUDART_INF_F)) return -1.f;
    const uint4 *row = reinterpret_cast<const uint4 *>(P.nbx + (size_t)sp * FUSED_NBX_K);
    uint4 e[FUSED_NBX_K / 4];
#pragma unroll
    for (int t = 0; t < (int)FUSED_NBX_K / 4; ++t) e[t] = __ldg(row + t);
    float best = ds, thr = tri_thr(ds, ds);
    uint32_t bi = sp;
    if (!(__uint_as_float(e[FUSED_NBX_K / 4 - 1].w & 0xFFFF0000u) > thr)) return -1.f;      // the walk must end inside the row
#pragma unroll
    for (int t = 0; t < (int)FUSED_NBX_K / 4; ++t)
    {
        const uint32_t c4[4] = { e[t].x, e[t].y, e[t].z, e[t].w };
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
            if (__uint_as_float(c4[u] & 0xFFFF0000u) > thr) { P.nn_o[gi] = bi; return best; }
            const uint32_t k = o + (c4[u] & 0xFFFFu);
            const pt8 x = ld_pt8(P.Xp, k);
            const float d = FAST ? dist6(q.lo, q.hi, x.lo, x.hi, fg, fp) : dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);
            ++ecnt;
            if (d < best || (d == best && k < bi)) { best = d; bi = k; thr = tri_thr(ds, best); }
        }
    }
    P.nn_o[gi] = bi;        // not reached: the last entry exceeds thr (checked above) and thr never grows
    return best;
}

// scan_reps for one point per lane group that also keeps a lower bound `sec` of the distance to every representative but
// the winner (stage-1 flavour): an evaluated distance enters exactly, a representative skipped by the
// partial-distance early-out enters with its geometric part pg (<= d by monotonic rounding).  s0 = the seed, evaluated
// first; when the loop meets it again it is the same representative and must not count as "another" one.
template <int S, bool FAST>
__device__ __forceinline__ void scan_reps_sec(const float4 *__restrict__ sRlo, const float4 *__restrict__ sRhi, uint32_t nr, uint32_t c,
                                              const pt8 &q, float &best, uint32_t &bi, float &sec, float fg, float fp, bool prune)
{
    const uint32_t s0 = bi;
    bool seeded;
    {
        const float4 rlo = sRlo[bi], rhi = sRhi[bi];
        const float d = FAST ? dist6(q.lo, q.hi, rlo, rhi, fg, fp) : dist8(q.lo, q.hi, rlo, rhi, fg, fp);
        seeded = d < CUDART_INF_F && prune;
        if (seeded) best = d;
        else { best = CUDART_INF_F; bi = c; }
    }
    sec = CUDART_INF_F;
#pragma unroll 2
    for (uint32_t r = c; r < nr; r += S)
    {
        const float4 rlo = sRlo[r];
        const float d0 = __fsub_rn(q.lo.x, rlo.x), d1 = __fsub_rn(q.lo.y, rlo.y), d2 = __fsub_rn(q.lo.z, rlo.z);
        float g = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
        if (!FAST) { const float d3 = __fsub_rn(q.lo.w, rlo.w); g = __fadd_rn(g, __fmul_rn(d3, d3)); }
        const float pg = __fmul_rn(fg, g);
        const bool pass = pg <= best;
        const bool other = !(seeded && r == s0);
        if (__any_sync(FULL_MASK, pass || !prune))
        {
            const float4 rhi = sRhi[r];
            const float d4 = __fsub_rn(q.hi.x, rhi.x), d5 = __fsub_rn(q.hi.y, rhi.y), d6 = __fsub_rn(q.hi.z, rhi.z);
            float p = __fadd_rn(__fadd_rn(__fmul_rn(d4, d4), __fmul_rn(d5, d5)), __fmul_rn(d6, d6));
            if (!FAST) { const float d7 = __fsub_rn(q.hi.w, rhi.w); p = __fadd_rn(p, __fmul_rn(d7, d7)); }
            const float d = __fadd_rn(pg, __fmul_rn(fp, p));
            if (other) sec = fminf(sec, fmaxf(d, best));
            if (d < best || (d == best && r < bi)) { best = d; bi = r; }
        }
        else if (other) sec = fminf(sec, pg);        // pg > best here, and d >= pg
    }
}

// exhaustive scan (seeded + early-out, scan_reps) of the chunk's points listed in fbl[0..nfb): SF lanes per point
template <int SF, bool SEARCH, bool SETTLE, bool ILP = false>
__device__ __forceinline__ void full_scan_pass(const PairPtrs &P, const float *X, const uint32_t *q_rep, const uint16_t *fbl, uint32_t nfb,
                                               uint32_t q0, uint32_t nr, const float4 *sRlo, const float4 *sRhi, uint32_t *keys,
                                               bool reps_w_const, const float4 &r0lo, const float4 &r0hi, const float4 &tq, const float4 &tt,
                                               float fg, float fp, bool prune, uint32_t &ecnt, const uint32_t k_now, const uint32_t nbx_m)
{
    const uint32_t tid = threadIdx.x, TPB = blockDim.x;
    float *lb1 = reinterpret_cast<float *>(P.nbx);            // [m] stage-1 runner-up bounds (settle flavour; nbx is free without nn_walk)
    uint32_t *tag1 = P.nbx + nbx_m;                           // [m] iteration the bound belongs to
    for (uint32_t t0 = 0; t0 < nfb; t0 += TPB / SF)
    {
        const uint32_t t = t0 + tid / SF, c = tid % SF;
        if (t0 + (tid & ~31u) / SF >= nfb) continue;          // no point for this warp in this pass (warp-uniform)
        const bool valid = t < nfb;
        const uint32_t l = fbl[valid ? t : 0u];
        const uint32_t gi = q0 + l;
        pt8 q[1];
        float best[1];
        uint32_t bi[1];
        q[0] = ld_pt8(X, gi);
        if (SEARCH) q[0].lo = transform_q_xyz(q[0].lo, tq, tt);
        const bool fastp = reps_w_const && (q[0].lo.w == r0lo.w) && (q[0].hi.w == r0hi.w);
        const bool warp_fast = __all_sync(FULL_MASK, fastp);
        bi[0] = min(__ldcg(q_rep + gi), nr - 1u);
        float sec = -1.f;
        if (SETTLE)
        {
            if (warp_fast) scan_reps_sec<SF, true>(sRlo, sRhi, nr, c, q[0], best[0], bi[0], sec, fg, fp, prune);
            else scan_reps_sec<SF, false>(sRlo, sRhi, nr, c, q[0], best[0], bi[0], sec, fg, fp, prune);
        }
        else if (ILP)
        {
            if (warp_fast) scan_reps_ilp<SF, true>(sRlo, sRhi, nr, c, q[0], best[0], bi[0], fg, fp, prune);
            else scan_reps_ilp<SF, false>(sRlo, sRhi, nr, c, q[0], best[0], bi[0], fg, fp, prune);
        }
        else if (warp_fast) scan_reps<SF, 1, true>(sRlo, sRhi, nr, c, q, best, bi, fg, fp, prune);
        else scan_reps<SF, 1, false>(sRlo, sRhi, nr, c, q, best, bi, fg, fp, prune);
        float b = best[0];
        uint32_t id = bi[0];
#pragma unroll
        for (int off = 1; off < SF; off <<= 1)
        {
            const float od = __shfl_xor_sync(FULL_MASK, b, off);
            const uint32_t oi = __shfl_xor_sync(FULL_MASK, id, off);
            if (SETTLE)
            {
                // two lane groups may hold the SAME representative as their best (the common seed): it is not its own runner-up
                const float os = __shfl_xor_sync(FULL_MASK, sec, off);
                sec = (oi == id) ? fminf(sec, os) : fminf(fminf(sec, os), fmaxf(b, od));
            }
            if (od < b || (od == b && oi < id)) { b = od; id = oi; }
        }
        if (valid && c == 0)
        {
            keys[l] = (b < CUDART_INF_F) ? id : 0u;
            ecnt += nr + 1u;
            if (SETTLE)
            {
                const bool usable = (b < CUDART_INF_F) && (sec > 1e-30f);
                lb1[gi] = usable ? __fmul_rd(__fsqrt_rd(sec), 0.999999f) : -1.f;
                tag1[gi] = k_now;
            }
        }
    }
}

// Body of kernel A for chunk `bx` of pair P (the CTA's shared memory starts at smem_a).  Called by k_assign_tri (one CTA per
// chunk) and by the persistent iteration kernel (k_icp_persistent), where the CTAs loop over the chunks.
// APERM: seed-grouped lane order of the pruned pass (batch engine); SETTLE: temporal pruning of stage 1 -- both compile-time, so
// that the latency-mode instantiation carries neither (with run-time flags it grew by 30 % and lost 0.5 us per iteration)
template <bool SEARCH, bool APERM, bool SETTLE>
__device__ __forceinline__ void assign_tri_body(const PairPtrs &P, const FusedCfg &cfg, const int tri_cfg, const uint32_t bx, float4 *smem_a)
{
    const uint32_t nr = cfg.nr, m = cfg.m, QB = cfg.QB, TPB = blockDim.x, K = cfg.K;
    float4 *sRlo = smem_a;                                       // [nr] xyz1 halves
    float4 *sRhi = sRlo + nr;                                    // [nr] rgb1 halves
    uint32_t *keys = reinterpret_cast<uint32_t *>(sRhi + nr);    // [QB]
    uint32_t *cnt = keys + QB;                                   // [nr]
    uint16_t *slc = reinterpret_cast<uint16_t *>(cnt + nr);      // [ceil(QB/32)][nr] (parallel ranking only)
    const uint32_t nsl = (QB + 31u) / 32u;
    const bool par_rank = cfg.par_rank != 0;
    uint32_t *fb_n = reinterpret_cast<uint32_t *>(slc + (par_rank ? ((nsl * nr + 1u) & ~1u) : 0u));   // [1] (+1 pad)
    uint16_t *fbl = reinterpret_cast<uint16_t *>(fb_n + 2);      // [2][QB] local indices of the points that need the full scan (second half: after the temporal filter)
    // the convergence flag is fetched now and tested after the first barrier (before any global write): its latency
    // overlaps the staging of the representatives instead of preceding it
    uint32_t done = 0u;
    if (SEARCH) done = __ldcg(&P.state->done);
    if (SEARCH) { PROF_STAMP(P, 0, 0, gtime_ns()); PROF_STAMP(P, 0, 1, (unsigned long long)clock64()); }
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    // (Measured, round 2: requesting the first trip's point / seed / pose before the staging loop, to put one memory latency
    // instead of two in front of the first distance, made the iteration 1 us SLOWER -- the staging loads queue behind them.)
    const float *X = SEARCH ? P.M : P.F;
    uint32_t *q_rep = SEARCH ? P.q_rep : P.rep_id;
    const uint32_t q0 = bx * QB;
    const uint32_t nq = min(QB, m - q0);
    const float4 r0lo = __ldg((const float4 *)P.reps), r0hi = __ldg((const float4 *)P.reps + 1);
    bool okw = finite_f(r0lo.w) && finite_f(r0hi.w);
    for (uint32_t i = tid; i < nr * 2u; i += TPB)
    {
        const float4 v = __ldg((const float4 *)P.reps + i);
        if (i & 1u) { sRhi[i >> 1] = v; okw = okw && (v.w == r0hi.w); }
        else { sRlo[i >> 1] = v; okw = okw && (v.w == r0lo.w); }
    }
    for (uint32_t i = tid; i < nr; i += TPB) cnt[i] = 0u;
    if (par_rank) for (uint32_t i = tid; i < (nsl * nr + 1u) / 2u; i += TPB) reinterpret_cast<uint32_t *>(slc)[i] = 0u;
    if (tid == 0) *fb_n = 0u;
    // The validity of the lane-order permutation must be ONE decision per CTA (CTA 0 of this launch re-tags lperm while later
    // CTAs start: threads of one CTA reading the tag at different times would mix the two lane orders and leave keys unwritten).
    __shared__ uint32_t s_perm_ok;
    if (APERM && tid == 0) s_perm_ok = (__ldcg(P.wconst + 12) == __ldcg(&P.state->k) + 1u) ? 1u : 0u;
    const bool reps_w_const = __syncthreads_and(okw) != 0;
    if (done) return;
    if (SEARCH) PROF_STAMP(P, 0, 2, (unsigned long long)clock64());

    float4 tq, tt;
    if (SEARCH) { tq = ICP_LOAD_T((const float4 *)P.T); tt = ICP_LOAD_T((const float4 *)P.T + 1); }
    const float fg = cfg.fg, fp = cfg.fp;
    const bool prune = fp >= 0.f;
    // stage-1 temporal pruning of the points that need the exhaustive scan (batch engine, metric weights in [0, 1])
    const bool settle1 = SEARCH && SETTLE && cfg.settle != 0 && cfg.nn_walk == 0 && fg >= 0.f && fg <= 1.f && fp >= 0.f && fp <= 1.f;
    const bool bounds_ok = settle1 && __ldcg(P.wconst + 13) != 0u;      // else: this iteration only records fresh bounds
    uint32_t k_now = 0u;
    const bool tri = tri_cfg != 0 && __ldcg(P.wconst + 1) != 0u;
    const uint2 *__restrict__ nbr = P.nbr;

    uint32_t ecnt = 0, ecnt2 = 0;            // stage-1 / stage-2 distance evaluations of this thread (reported when P.evals is set)
    const bool walk2 = SEARCH && tri && cfg.nn_walk != 0;
    const bool fx_const = SEARCH && __ldcg(P.wconst) != 0u;     // every fixed point carries the constant homogeneous lanes
    if (SEARCH && cfg.nn_walk != 0 && !walk2)
        for (uint32_t l = tid; l < nq; l += TPB) P.nnd[q0 + l] = -1.f;
    // Lane order of the pruned pass: the chunk's points grouped by the representative they had last iteration (local
    // permutation written by the previous iteration's rank pass; any bijection is correct, this one makes the lanes of a warp
    // walk the neighbourhood of the same seed: uniform walk lengths, broadcast shared-memory reads).  Valid once this
    // registration has completed an iteration (state->k > 0: kernel D counts them, reset / k_batch_reset clear it).
    uint16_t *lperm = reinterpret_cast<uint16_t *>(P.nbx + 2u * (size_t)m);      // [m] u16, third region of nbx (free without nn_walk)
    const bool aperm = SEARCH && APERM && cfg.aperm != 0 && cfg.nn_walk == 0 && par_rank;
    const bool use_perm = APERM && aperm && s_perm_ok != 0u;
    // stage-1 temporal pruning: lb1[i] = proven lower bound of sqrt(D) between point i and every representative
    // but its last winner, refreshed for EVERY point in EVERY iteration (settled: lowered by the motion; walked: from the walk's
    // runner-up and stop entry; exhaustively scanned: from the scan's runner-up), so it is never stale while wconst[13] holds
    float *lb1 = reinterpret_cast<float *>(P.nbx);
    float4 pq, pt;
    if (settle1) { pq = __ldcg((const float4 *)(P.wconst + 4)); pt = __ldcg((const float4 *)(P.wconst + 4) + 1); k_now = __ldcg(&P.state->k); }
    // ---- pruned pass: one point per lane ----
    // (measured: prefetching the later trips' points / neighbour rows up front costs more than it hides: 0.218 -> 0.234 ms)
    for (uint32_t l0 = 0; l0 < nq; l0 += TPB)
    {
        const bool valid = l0 + tid < nq;
        const uint32_t l = (APERM && valid && use_perm) ? (uint32_t)__ldcg(lperm + q0 + l0 + tid) : l0 + tid;
        const uint32_t gi = q0 + (valid ? l : 0u);
        pt8 q = ld_pt8(X, gi);
        const float4 mlo = q.lo;
        const float lbv = (SEARCH && SETTLE && settle1) ? __ldcg(lb1 + gi) : -1.f;
        if (SEARCH) q.lo = transform_q_xyz(q.lo, tq, tt);
        const bool fastp = reps_w_const && (q.lo.w == r0lo.w) && (q.hi.w == r0hi.w);
        const bool warp_fast = __all_sync(FULL_MASK, fastp);
        if (!SEARCH && !warp_fast && lane == 0) *P.wconst = 0u;
        const uint32_t s = min(__ldcg(q_rep + gi), nr - 1u);
        float ds;
        ecnt += valid ? 1u : 0u;
        if (warp_fast) ds = dist6(q.lo, q.hi, sRlo[s], sRhi[s], fg, fp);
        else ds = dist8(q.lo, q.hi, sRlo[s], sRhi[s], fg, fp);
        bool settled = false;
        if (SEARCH && SETTLE && settle1 && valid && bounds_ok && lbv > 0.f)
        {
            // the point moved by at most delta in the metric space since the bound was recorded (same test as in kernel C')
            const float4 qp = transform_q_xyz(mlo, pq, pt);
            const float dx = fmaxf(fabsf(__fsub_ru(q.lo.x, qp.x)), fabsf(__fsub_rd(q.lo.x, qp.x)));
            const float dy = fmaxf(fabsf(__fsub_ru(q.lo.y, qp.y)), fabsf(__fsub_rd(q.lo.y, qp.y)));
            const float dz = fmaxf(fabsf(__fsub_ru(q.lo.z, qp.z)), fabsf(__fsub_rd(q.lo.z, qp.z)));
            const float s2 = __fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz));
            const float lbn = __fsub_rd(lbv, __fsqrt_ru(__fmul_ru(fg, s2)));
            if (lbn > 0.f && __fmul_rd(__fmul_rd(lbn, lbn), 0.99999f) > __fadd_ru(ds, 1e-30f))
            {
                settled = true;                 // every other representative's computed distance exceeds ds: s wins again
                keys[l] = s;
                lb1[gi] = lbn;
            }
        }
        const uint2 *row = nbr + (size_t)s * K;
        uint32_t r = 0xFFFFFFFFu;
        float sec = CUDART_INF_F, dstop = -1.f;
        if (tri && valid && !settled && (ds < CUDART_INF_F))
        {
            if (SEARCH && SETTLE && settle1)
                r = warp_fast ? tri_walk<true, true>(row, K, q, sRlo, sRhi, fg, fp, ds, s, ecnt, sec, dstop)
                              : tri_walk<false, true>(row, K, q, sRlo, sRhi, fg, fp, ds, s, ecnt, sec, dstop);
            else
                r = warp_fast ? tri_walk<true, false>(row, K, q, sRlo, sRhi, fg, fp, ds, s, ecnt, sec, dstop)
                              : tri_walk<false, false>(row, K, q, sRlo, sRhi, fg, fp, ds, s, ecnt, sec, dstop);
        }
        if (settled) { }
        else if (r != 0xFFFFFFFFu)
        {
            keys[l] = r;
            if (SEARCH && SETTLE && settle1)
            {
                // lower bound of sqrt(D) to every representative but r: the evaluated ones through the runner-up, the others
                // through the triangle inequality at the stop entry; every rounding against the bound
                float lb = -1.f;
                if (dstop > 0.f && sec > 1e-30f)
                {
                    const float lb_un = __fsub_rd(__fmul_rd(__fsqrt_rd(dstop), 0.999999f), __fmul_ru(__fsqrt_ru(ds), 1.000001f));
                    const float lb_ev = (sec < CUDART_INF_F) ? __fmul_rd(__fsqrt_rd(sec), 0.999999f) : CUDART_INF_F;
                    lb = fminf(lb_un, lb_ev);
                    if (!(lb > 0.f) || !(lb < CUDART_INF_F)) lb = -1.f;
                }
                lb1[gi] = lb;
            }
            if (SEARCH && walk2)
                P.nnd[gi] = (warp_fast && fx_const) ? nn_walk<true>(P, q, r, gi, fg, fp, ecnt2) : nn_walk<false>(P, q, r, gi, fg, fp, ecnt2);
        }
        else if (valid)
        {
            fbl[atomicAdd(fb_n, 1u)] = (uint16_t)l;
            if (SEARCH && walk2) P.nnd[gi] = -1.f;          // representative not known yet: kernel C scans its list
        }
    }
    __syncthreads();
    if (SEARCH) PROF_STAMP(P, 0, 3, (unsigned long long)clock64());
    // ---- full scan of the points the bound could not settle: SF lanes per point (8 in batch mode: fewer instructions;
    //      32 in latency mode: a 4x shorter dependent chain per point) ----
    uint32_t nfb = *fb_n;
    if (SEARCH && P.prof && tid == 0) { atomicAdd(P.prof + 44, (unsigned long long)nfb); atomicAdd(P.prof + 45, 1ull); }
    // lanes per point of the exhaustive scan: 8 (fewest instructions) when the list is long; with few points left (late
    // iterations: most outliers settle) 32 lanes per point make the dependent chain of the one remaining pass 4x shorter
    if (SEARCH && SETTLE && settle1 && nfb * 32u <= TPB) full_scan_pass<32, SEARCH, SEARCH && SETTLE>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    else if (SEARCH && SETTLE && settle1) full_scan_pass<TRI_S, SEARCH, SEARCH && SETTLE>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    // the independent-evaluation flavours exist only in the instantiation without the temporal pruning (latency mode): compiled
    // into the batch-engine instance they cost kernel A 2 % (0.2097 -> 0.2140 ms per 256-pair launch) without ever running there
    else if (!SETTLE && cfg.SF == 32) full_scan_pass<32, SEARCH, false, true>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    else if (!SETTLE && cfg.SF == 16) full_scan_pass<16, SEARCH, false, true>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    else if (!SETTLE && cfg.SF == 9) full_scan_pass<TRI_S, SEARCH, false, true>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    else if (cfg.SF == 32) full_scan_pass<32, SEARCH, false>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    else full_scan_pass<TRI_S, SEARCH, false>(P, X, q_rep, fbl, nfb, q0, nr, sRlo, sRhi, keys, reps_w_const, r0lo, r0hi, tq, tt, fg, fp, prune, ecnt, k_now, m);
    if (SEARCH && P.evals)
    {
        unsigned long long e = ecnt, e2 = ecnt2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { e += __shfl_down_sync(FULL_MASK, e, d); e2 += __shfl_down_sync(FULL_MASK, e2, d); }
        if (lane == 0 && e) atomicAdd(P.evals + 2, e);
        if (lane == 0 && e2) atomicAdd(P.evals + 3, e2);
    }
    if (SEARCH) PROF_STAMP(P, 0, 4, (unsigned long long)clock64());
    // scratch of the rank pass: the staged representatives are dead by now (chunk_rank_store starts with a barrier)
    if (APERM) chunk_rank_store(P, cfg, keys, cnt, slc, q_rep, q0, nq, bx, aperm ? lperm : nullptr, reinterpret_cast<uint32_t *>(sRhi));
    else chunk_rank_store(P, cfg, keys, cnt, slc, q_rep, q0, nq, bx);
    if (SEARCH) { PROF_STAMP(P, 0, 5, (unsigned long long)clock64()); PROF_STAMP(P, 0, 6, gtime_ns()); PROF_END_ALL(P, 0); }
}

template <bool SEARCH, bool APERM, bool SETTLE>
__global__ void __launch_bounds__(512, 2) k_assign_tri(const PairPtrs *__restrict__ table, const FusedCfg cfg, const int tri_cfg)
{
    extern __shared__ float4 smem_a[];
    pdl_wait(); pdl_trigger();
    const PairPtrs P = table[blockIdx.y];
    assign_tri_body<SEARCH, APERM, SETTLE>(P, cfg, tri_cfg, blockIdx.x, smem_a);
}

// =================================================================================================
// B: per representative (column), exclusive prefix of the chunk histograms over the chunks (rows);
// column totals = list sizes.  CTA = 32 columns x 8 row-slabs.
// =================================================================================================
#define COLSCAN_MAXPER 8
template <bool SEARCH>
__device__ __forceinline__ void colscan_body(const PairPtrs &P, const FusedCfg &cfg, const uint32_t bx)
{
    __shared__ uint32_t ws[32][33];
    uint32_t done = 0u;
    if (SEARCH) done = __ldcg(&P.state->done);              // tested after the barrier, before the first global write
    if (SEARCH) PROF_STAMP(P, 1, 0, gtime_ns());
    const uint32_t nr = cfg.nr, nb = cfg.nbA;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t r = bx * 32u + lane;
    const uint32_t nw = blockDim.x >> 5;                    // warps of the CTA (<= 32): one slab of rows each
    const uint32_t per = (nb + nw - 1u) / nw;               // rows per warp (slab)
    const uint32_t row0 = w * per, row1 = min(nb, row0 + per);
    uint32_t sum = 0;
    uint32_t v[COLSCAN_MAXPER];
    const bool inreg = per <= COLSCAN_MAXPER;
    if (r < nr)
    {
        if (inreg)
        {
#pragma unroll
            for (uint32_t j = 0; j < COLSCAN_MAXPER; ++j)
            {
                const uint32_t row = row0 + j;
                v[j] = (row < row1) ? __ldcg(P.H + (size_t)row * nr + r) : 0u;
            }
#pragma unroll
            for (uint32_t j = 0; j < COLSCAN_MAXPER; ++j) sum += v[j];
        }
        else
        {
#pragma unroll 4
            for (uint32_t row = row0; row < row1; ++row) sum += __ldcg(P.H + (size_t)row * nr + r);
        }
    }
    ws[w][lane] = sum;
    __syncthreads();
    if (done) return;
    uint32_t base = 0, total = 0;
    for (uint32_t w2 = 0; w2 < nw; ++w2) { const uint32_t t = ws[w2][lane]; if (w2 < w) base += t; total += t; }
    if (r < nr)
    {
        uint32_t run = base;
        if (inreg)
        {
#pragma unroll
            for (uint32_t j = 0; j < COLSCAN_MAXPER; ++j)
            {
                const uint32_t row = row0 + j;
                if (row < row1) { P.H[(size_t)row * nr + r] = run; run += v[j]; }
            }
        }
        else
        {
            for (uint32_t row = row0; row < row1; ++row)
            {
                const uint32_t t = __ldcg(P.H + (size_t)row * nr + r);
                P.H[(size_t)row * nr + r] = run;
                run += t;
            }
        }
        if (w == 0) (SEARCH ? P.Nq : P.N)[r] = total;
    }
}

template <bool SEARCH>
__global__ void __launch_bounds__(1024) k_colscan(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    pdl_wait(); pdl_trigger();
    const PairPtrs P = table[blockIdx.y];
    colscan_body<SEARCH>(P, cfg, blockIdx.x);
}

// exclusive scan of cnt[0..nr) into shared memory by the whole CTA (any block size that is a multiple of 32, <= 1024)
__device__ __forceinline__ void cta_exscan_to_smem(const uint32_t *__restrict__ cnt, uint32_t nr, uint32_t *out_s, uint32_t *warp_tot)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const uint32_t per = (nr + nthreads - 1u) / nthreads;
    const uint32_t b0 = tid * per;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; ++j) if (b0 + j < nr) sum += __ldcg(cnt + b0 + j);
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= (uint32_t)d) inc += v;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t w2 = 0; w2 < w && w2 < nwarps; ++w2) wbase += warp_tot[w2];
    uint32_t run = wbase + inc - sum;
    for (uint32_t j = 0; j < per; ++j)
        if (b0 + j < nr) { out_s[b0 + j] = run; run += __ldcg(cnt + b0 + j); }
    __syncthreads();
}

// build only: scatter the fixed set into its lists (X_p), perm, list offsets O
__global__ void __launch_bounds__(256) k_build_scatter(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ uint32_t smem_o[];
    __shared__ uint32_t warp_tot[32];
    const PairPtrs P = table[blockIdx.y];
    const uint32_t nr = cfg.nr, m = cfg.m;
    cta_exscan_to_smem(P.N, nr, smem_o, warp_tot);
    if (blockIdx.x == 0) for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x) P.O[r] = smem_o[r];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t k = __ldcg(P.rep_id + i);
    const uint32_t pos = smem_o[k] + __ldcg(P.H + (size_t)(i / cfg.QB) * nr + k) + __ldcg(P.lrank + i);
    P.perm[pos] = i;
    P.q_rep[i] = k;                          // seeds of the first search iteration: the moving point starts near its fixed twin
    P.nn_o[i] = pos;
    P.nnd[i] = -1.f;                         // sorted flavour: no proven runner-up bound yet (first search scans every list)
    P.nn2[i] = make_uint2(0xBF800000u, pos);
    if (cfg.settle && !cfg.nn_walk) { P.nbx[i] = 0xBF800000u; P.nbx[m + i] = 0x7FFFFFFFu; }     // stage-1 bounds: none
    st_pt8(P.Xp, pos, ld_pt8(P.F, i));
}

// =================================================================================================
// C: sorted position of every query + stage-2 list scan + weight + scatter into the sorted SoA arrays.
// =================================================================================================
template <int L>
__global__ void __launch_bounds__(256) k_search(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ uint32_t smem_o[];
    __shared__ uint32_t warp_tot[32];
    const PairPtrs P = table[blockIdx.y];
    if (P.state->done) return;
    const uint32_t nr = cfg.nr, m = cfg.m;
    cta_exscan_to_smem(P.Nq, nr, smem_o, warp_tot);
    if (blockIdx.x == 0) for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x) P.Oq[r] = smem_o[r];
    // L adjacent lanes share one query and scan list positions c, c+L, c+2L, ... (coalesced 32*L-byte reads);
    // ordered argmin merge (lower distance, then lower list position) == sequential strict-'<' scan.
    // A CTA owns cfg.QC consecutive queries and walks them 256/L at a time (the prologue is paid once).
    constexpr uint32_t QPC = 256 / L;
    const uint32_t c = threadIdx.x % L;
    const float4 tq = __ldg((const float4 *)P.T), tt = __ldg((const float4 *)P.T + 1);
    const float fg = cfg.fg, fp = cfg.fp;
    unsigned long long e_cnt = 0;
    for (uint32_t pass = 0; pass < cfg.QC; pass += QPC)
    {
        const uint32_t i = blockIdx.x * cfg.QC + pass + threadIdx.x / L;
        const bool valid = i < m && (pass + threadIdx.x / L) < cfg.QC;
        uint32_t cnt = 0, pos = 0, o = 0;
        pt8 q;
        q.lo = make_float4(0.f, 0.f, 0.f, 0.f); q.hi = q.lo;
        float best = CUDART_INF_F;
        uint32_t bi = 0;
        if (valid)
        {
            const uint32_t r = __ldcg(P.q_rep + i);
            pos = smem_o[r] + __ldcg(P.H + (size_t)(i / cfg.QB) * nr + r) + __ldcg(P.lrank + i);
            q = ld_pt8(P.M, i);
            q.lo = transform_q_xyz(q.lo, tq, tt);
            o = __ldg(P.O + r);
            cnt = __ldg(P.N + r);
            bi = o;
#pragma unroll 2
            for (uint32_t k = o + c; k < o + cnt; k += L)
            {
                const pt8 x = ld_pt8(P.Xp, k);
                const float d = dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);
                if (d < best) { best = d; bi = k; }
            }
        }
#pragma unroll
        for (int off = 1; off < L; off <<= 1)
        {
            const float od = __shfl_xor_sync(FULL_MASK, best, off);
            const uint32_t oi = __shfl_xor_sync(FULL_MASK, bi, off);
            if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
        }
        if (valid && c == 0)
        {
            if (best == CUDART_INF_F) bi = o;           // nothing compared less than +inf: the sequential scan keeps the list head
            if (cnt == 0) bi = o ? o - 1u : 0u;
            if (bi >= m) bi = m - 1u;
            const float4 nn = __ldg((const float4 *)P.Xp + (size_t)bi * 2);
            P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, best));
            P.fxyz[pos] = nn.x; P.fxyz[(size_t)m + pos] = nn.y; P.fxyz[(size_t)2 * m + pos] = nn.z;
            P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
            icp_dist_id di; di.dist = best; di.id = bi;
            P.NNID[pos] = di;
            P.qperm[pos] = i;
            e_cnt += cnt;
        }
    }
    if (P.evals)
    {
        unsigned long long e = e_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(FULL_MASK, e, d);
        if ((threadIdx.x & 31u) == 0 && e) atomicAdd(P.evals + 1, e);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.evals, (unsigned long long)m * nr);
    }
}

// exclusive scan of a shared-memory array (n entries) into another shared-memory array; returns the total in every thread
__device__ __forceinline__ uint32_t cta_exscan_smem(const uint32_t *in_s, uint32_t n, uint32_t *out_s, uint32_t *warp_tot)
{
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5, nthreads = blockDim.x, nwarps = nthreads >> 5;
    const uint32_t per = (n + nthreads - 1u) / nthreads;
    const uint32_t b0 = tid * per;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; ++j) if (b0 + j < n) sum += in_s[b0 + j];
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= (uint32_t)d) inc += v;
    }
    __syncthreads();                       // warp_tot may still be read by a previous scan
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    // only the first ceil(n / per) threads hold elements: the warps behind them contribute nothing
    const uint32_t nact = min(nwarps, ((n + per - 1u) / per + 31u) >> 5);
    uint32_t wbase = 0, total = 0;
    for (uint32_t w2 = 0; w2 < nact; ++w2) { const uint32_t t = warp_tot[w2]; if (w2 < w) wbase += t; total += t; }
    uint32_t run = wbase + inc - sum;
    for (uint32_t j = 0; j < per; ++j)
        if (b0 + j < n) { const uint32_t v = in_s[b0 + j]; out_s[b0 + j] = run; run += v; }
    __syncthreads();
    return total;
}

// =================================================================================================
// C (grouped): the CTA owns CC consecutive A-chunks.  Its queries are transformed once and staged in shared
// memory grouped by representative (their stable local order is already known: chunk histogram prefixes H +
// in-chunk ranks lrank), and the groups are cut into work items of <= QI queries that the warps pull from a
// shared counter.  Inside an item the lanes are w = pow2ceil(#queries) queries x P = 32/w list phases; the
// representative's list streams through a per-warp shared-memory tile of 32 points (coalesced 1 KB loads, the
// next tile is in flight in registers while the current one is scanned), and every lane reads the tile by
// broadcast LDS.128: no divergence on the list length, no per-lane global latency in the distance loop.  The
// ordered (distance, position) merge over the P phases reproduces the sequential strict-'<' scan, so the
// outputs are identical to k_search<L>.
// =================================================================================================
#ifndef GROUPED_WARPS
#define GROUPED_WARPS 8
#endif
#ifndef SCAN_FULL_UNROLL
#define SCAN_FULL_UNROLL 32      // evaluations per loop trip of the full-tile scans
#endif
#define ICP_DO_PRAGMA(x) _Pragma(#x)
#define ICP_UNROLL(n) ICP_DO_PRAGMA(unroll n)
#ifndef GROUPED_MINB
#define GROUPED_MINB 3
#endif
struct GroupedSmem
{
    float4 *qlo, *qhi;          // [QC] transformed queries, indexed by local query
    float4 *tile;               // [GROUPED_WARPS][64] per-warp list tile: 32 xyz1 halves, then 32 rgb1 halves
    uint32_t *sOq, *cnt, *offC, *ibase, *nsl, *sO, *sN;   // [nr] each
    uint32_t *items;            // [nr + QC/QI + 1]
    uint32_t *spos, *sidx;      // [QC] spos[l]: global sorted position of local query l; sidx[slot]: local query of a group slot
    uint32_t *rs;               // [QC] representative | (slot inside its group << 16) of local query l; 0xFFFFFFFF = already matched
};
__host__ __device__ static inline size_t grouped_carve(GroupedSmem *g, void *base, uint32_t nr, uint32_t QC, uint32_t QI, uint32_t warps = GROUPED_WARPS)
{
    char *p = (char *)base;
    size_t off = 0;
    if (g) g->qlo = (float4 *)(p + off); off += (size_t)QC * 16;
    if (g) g->qhi = (float4 *)(p + off); off += (size_t)QC * 16;
    if (g) g->tile = (float4 *)(p + off); off += (size_t)warps * 64 * 16;
    uint32_t **arr[7] = { g ? &g->sOq : nullptr, g ? &g->cnt : nullptr, g ? &g->offC : nullptr, g ? &g->ibase : nullptr,
                          g ? &g->nsl : nullptr, g ? &g->sO : nullptr, g ? &g->sN : nullptr };
    for (int i = 0; i < 7; ++i) { if (g) *arr[i] = (uint32_t *)(p + off); off += (size_t)nr * 4; }
    if (g) g->items = (uint32_t *)(p + off); off += (size_t)(nr + QC / QI + 1) * 4;
    if (g) g->spos = (uint32_t *)(p + off); off += (size_t)QC * 4;
    if (g) g->sidx = (uint32_t *)(p + off); off += (size_t)QC * 4;
    if (g) g->rs = (uint32_t *)(p + off); off += (size_t)QC * 4;
    return off + 16;
}

template <bool FAST>
__device__ __forceinline__ void scan_tile(const float4 *tlo, const float4 *thi, uint32_t tl, uint32_t ph, uint32_t Pn, uint32_t kbase,
                                          const pt8 &q, float fg, float fp, float &best, uint32_t &bi)
{
#pragma unroll 4
    for (uint32_t k = ph; k < tl; k += Pn)
    {
        const float4 xlo = tlo[k], xhi = thi[k];
        const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        if (d < best) { best = d; bi = kbase + k; }
    }
}

// one lane = one query (Pn == 1) against a full 32-point tile: fully unrolled, the tile is addressed with immediate offsets
template <bool FAST>
__device__ __forceinline__ void scan_tile_full(const float4 *tlo, const float4 *thi, uint32_t kbase,
                                               const pt8 &q, float fg, float fp, float &best, uint32_t &bi)
{
    uint32_t bk = 0xFFFFFFFFu;
ICP_UNROLL(SCAN_FULL_UNROLL)
    for (uint32_t k = 0; k < 32u; ++k)
    {
        const float4 xlo = tlo[k], xhi = thi[k];
        const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        if (d < best) { best = d; bk = k; }
    }
    if (bk != 0xFFFFFFFFu) bi = kbase + bk;
}

// runner-up tracking flavours of the tile scans (defined with the sorted kernel below)
template <bool FAST>
__device__ __forceinline__ void scan_tile_sec(const float4 *tlo, const float4 *thi, uint32_t tl, uint32_t ph, uint32_t Pn, uint32_t kbase,
                                              const pt8 &q, float fg, float fp, float &best, uint32_t &bi, float &sec);
template <bool FAST>
__device__ __forceinline__ void scan_tile_full_sec(const float4 *tlo, const float4 *thi, uint32_t kbase,
                                                   const pt8 &q, float fg, float fp, float &best, uint32_t &bi, float &sec);

// Body of the grouped kernel C for the `bx`-th run of QG consecutive queries (any CTA size that is a multiple of 32: one list
// tile per warp).  Called by k_search_grouped and by the persistent iteration kernel.
// compile-time: SETTLE (the latency-mode instantiation carries none of the temporal-pruning code) and the number of warps
// (the shared-memory carve-up folds into immediate offsets: with a run-time warp count the kernel lost 1.4 us per iteration)
template <bool SETTLE, int WARPS>
__device__ __forceinline__ void search_grouped_body(const PairPtrs &P, const FusedCfg &cfg, const uint32_t bx, float4 *smem_g4)
{
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_ctr;
    const uint32_t done = __ldcg(&P.state->done);           // tested after the first scan's barriers, before the first global write
    PROF_STAMP(P, 2, 0, gtime_ns()); PROF_STAMP(P, 2, 1, (unsigned long long)clock64());
    const uint32_t nr = cfg.nr, m = cfg.m, QB = cfg.QB, QI = cfg.QI;
    const uint32_t QC = cfg.QG;
    GroupedSmem G;
    grouped_carve(&G, smem_g4, nr, QC, QI, WARPS);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t q0 = bx * QC, nq_cta = min(QC, m - q0);
    // Image-ordered clouds (one large registration): the CTA owns a ctile_w x (QC / ctile_w) patch of the lm_w x lm_h grid instead of
    // QC consecutive points.  A run of 512 consecutive pixels of a 640-wide frame crosses 26 representative cells (20 x 15 pixels at
    // 1024 representatives) with ~20 queries each -- work items of 20 queries on 32 lanes; a 32 x 16 patch touches ~5 cells with
    // ~100 queries each.  Results go to the same sorted positions either way.
    const uint32_t tw = cfg.ctile_w;
    const uint32_t tiles_x = tw ? cfg.lm_w / tw : 1u;
    const uint32_t tbase = tw ? (bx / tiles_x) * (QC / tw) * cfg.lm_w + (bx % tiles_x) * tw : q0;
    auto qidx = [&](uint32_t l) -> uint32_t { return tw ? tbase + (l / tw) * cfg.lm_w + (l % tw) : q0 + l; };

    cta_exscan_to_smem(P.Nq, nr, G.sOq, warp_tot);
    if (done) return;
    PROF_STAMP(P, 2, 2, (unsigned long long)clock64());
    if (bx == 0) for (uint32_t r = tid; r < nr; r += blockDim.x) P.Oq[r] = G.sOq[r];
    for (uint32_t r = tid; r < nr; r += blockDim.x)
    {
        G.cnt[r] = 0u;
        G.sO[r] = __ldg(P.O + r);
        G.sN[r] = __ldg(P.N + r);
    }
    if (tid == 0) s_ctr = 0;
    __syncthreads();
    const float4 tq = ICP_LOAD_T((const float4 *)P.T), tt = ICP_LOAD_T((const float4 *)P.T + 1);
    // dist6 shortcut (see k_assign): every fixed point carries the homogeneous lanes of representative 0
    const float w_lo = __ldg(P.reps + 3), w_hi = __ldg(P.reps + 7);
    bool fast = __ldcg(P.wconst) != 0u;
    const bool walked = cfg.nn_walk != 0;
    // exact temporal pruning of stage 2, as in k_search_sorted: here per query of the chunk, in original order
    const bool settle = SETTLE && cfg.settle != 0 && !walked;
    const bool bounds_ok = settle && __ldcg(P.wconst + 13) != 0u;
    const float4 pq = __ldcg((const float4 *)(P.wconst + 4)), pt = __ldcg((const float4 *)(P.wconst + 4) + 1);
    const float fg = cfg.fg, fp = cfg.fp;
    unsigned long long e_cnt = 0, x_cnt = 0;
    // pass 1: sorted position of every query of the CTA.  Queries kernel A already matched (pruned walk from last
    // iteration's neighbour) are finished here; the others are counted per representative and parked in shared memory.
    for (uint32_t l = tid; l < nq_cta; l += blockDim.x)
    {
        const uint32_t i = qidx(l), c = i / QB;
        const uint32_t r = __ldcg(P.q_rep + i);
        const uint32_t h = __ldcg(P.H + (size_t)c * nr + r), lr = __ldcg(P.lrank + i);
        float nd = walked ? __ldcg(P.nnd + i) : -1.f;
        pt8 q = ld_pt8(P.M, i);
        const float4 mlo = q.lo;
        q.lo = transform_q_xyz(q.lo, tq, tt);
        const uint32_t pos = G.sOq[r] + h + lr;
        if (settle)
        {
            const float lbv = __ldcg(P.nnd + i);
            const uint32_t nno = __ldcg(P.nn_o + i);
            if (bounds_ok && lbv > 0.f && (nno - G.sO[r]) < G.sN[r])     // same representative as when x* was found
            {
                const float4 qp = transform_q_xyz(mlo, pq, pt);
                const float dx = fmaxf(fabsf(__fsub_ru(q.lo.x, qp.x)), fabsf(__fsub_rd(q.lo.x, qp.x)));
                const float dy = fmaxf(fabsf(__fsub_ru(q.lo.y, qp.y)), fabsf(__fsub_rd(q.lo.y, qp.y)));
                const float dz = fmaxf(fabsf(__fsub_ru(q.lo.z, qp.z)), fabsf(__fsub_rd(q.lo.z, qp.z)));
                const float s2 = __fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz));
                const float lbn = __fsub_rd(lbv, __fsqrt_ru(__fmul_ru(fg, s2)));
                const pt8 x = ld_pt8(P.Xp, nno);
                const float d = dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);     // == dist6 bit for bit whenever dist6 applies
                if (lbn > 0.f && __fmul_rd(__fmul_rd(lbn, lbn), 0.99999f) > __fadd_ru(d, 1e-30f))
                {
                    nd = d;                     // settled: finished below like a query the walk matched (nn_o[i] is x*)
                    P.nnd[i] = lbn;
                    x_cnt += 1u;
                }
            }
        }
        if (nd >= 0.f)
        {
            const uint32_t bi = __ldcg(P.nn_o + i);
            const float4 nn = __ldg((const float4 *)P.Xp + (size_t)bi * 2);
            P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, nd));
            P.fxyz[pos] = nn.x; P.fxyz[(size_t)m + pos] = nn.y; P.fxyz[(size_t)2 * m + pos] = nn.z;
            P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
            icp_dist_id di; di.dist = nd; di.id = bi;
            P.NNID[pos] = di;
            P.qperm[pos] = i;
            G.rs[l] = 0xFFFFFFFFu;
            e_cnt += G.sN[r];
        }
        else
        {
            fast = fast && (q.lo.w == w_lo) && (q.hi.w == w_hi);
            const uint32_t slot = atomicAdd(&G.cnt[r], 1u);
            G.qlo[l] = q.lo; G.qhi[l] = q.hi;
            G.spos[l] = pos;
            G.rs[l] = r | (slot << 16);
        }
    }
    fast = __syncthreads_and(fast) != 0;
    PROF_STAMP(P, 2, 3, (unsigned long long)clock64());
    for (uint32_t r = tid; r < nr; r += blockDim.x) G.nsl[r] = (G.cnt[r] + QI - 1u) / QI;
    __syncthreads();
    cta_exscan_smem(G.cnt, nr, G.offC, warp_tot);
    const uint32_t nitems = cta_exscan_smem(G.nsl, nr, G.ibase, warp_tot);
    for (uint32_t r = tid; r < nr; r += blockDim.x)
        for (uint32_t s = 0; s < G.nsl[r]; ++s) G.items[G.ibase[r] + s] = r | (s << 16);
    // pass 2: the parked queries take their slot in their representative's group (the order inside a group is irrelevant:
    // every result goes to its own sorted position)
    for (uint32_t l = tid; l < nq_cta; l += blockDim.x)
    {
        const uint32_t v = G.rs[l];
        if (v != 0xFFFFFFFFu) G.sidx[G.offC[v & 0xFFFFu] + (v >> 16)] = l;
    }
    __syncthreads();
    PROF_STAMP(P, 2, 4, (unsigned long long)clock64());

    float4 *tlo = G.tile + warp * 64u, *thi = tlo + 32;
    while (true)
    {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(&s_ctr, 1u);
        it = __shfl_sync(FULL_MASK, it, 0);
        if (it >= nitems) break;
        const uint32_t item = G.items[it];
        const uint32_t r = item & 0xFFFFu, sl = item >> 16;
        const uint32_t nq = min(QI, G.cnt[r] - sl * QI);
        const uint32_t lw = nq > 1u ? 32u - (uint32_t)__clz(nq - 1u) : 0u;
        const uint32_t w = 1u << lw;                             // queries (padded to a power of two) ...
        const uint32_t Pn = 32u >> lw;                           // ... x list phases
        const uint32_t ql = lane & (w - 1u), ph = lane >> lw;
        const bool valid = ql < nq;
        const uint32_t lq = G.sidx[G.offC[r] + sl * QI + (valid ? ql : 0u)];       // local query of this lane
        pt8 q; q.lo = G.qlo[lq]; q.hi = G.qhi[lq];
        const uint32_t o = G.sO[r], len = G.sN[r];
        float best = CUDART_INF_F, sec = CUDART_INF_F;
        uint32_t bi = o;
        pt8 nx;
        if (lane < len) nx = ld_pt8(P.Xp, o + lane);
        for (uint32_t t0 = 0; t0 < len; t0 += 32u)
        {
            const uint32_t tl = min(32u, len - t0);
            __syncwarp();
            if (lane < tl) { tlo[lane] = nx.lo; thi[lane] = nx.hi; }
            __syncwarp();
            if (t0 + 32u + lane < len) nx = ld_pt8(P.Xp, o + t0 + 32u + lane);
            if (settle)
            {
                if (Pn == 1u && tl == 32u)
                {
                    if (fast) scan_tile_full_sec<true>(tlo, thi, o + t0, q, fg, fp, best, bi, sec);
                    else scan_tile_full_sec<false>(tlo, thi, o + t0, q, fg, fp, best, bi, sec);
                }
                else if (fast) scan_tile_sec<true>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi, sec);
                else scan_tile_sec<false>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi, sec);
            }
            else if (Pn == 1u && tl == 32u)
            {
                if (fast) scan_tile_full<true>(tlo, thi, o + t0, q, fg, fp, best, bi);
                else scan_tile_full<false>(tlo, thi, o + t0, q, fg, fp, best, bi);
            }
            else if (fast) scan_tile<true>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi);
            else scan_tile<false>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi);
        }
        for (uint32_t off = w; off < 32u; off <<= 1)
        {
            const float od = __shfl_xor_sync(FULL_MASK, best, off);
            const uint32_t oi = __shfl_xor_sync(FULL_MASK, bi, off);
            if (settle)
            {
                const float os = __shfl_xor_sync(FULL_MASK, sec, off);
                sec = fminf(fminf(sec, os), fmaxf(best, od));        // runner-up of the union of the two phases
            }
            if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
        }
        if (valid && ph == 0)
        {
            if (best == CUDART_INF_F) bi = o;           // nothing compared less than +inf: the sequential scan keeps the list head
            if (len == 0) bi = o ? o - 1u : 0u;
            if (bi >= m) bi = m - 1u;
            const uint32_t pos = G.spos[lq];
            const float4 nn = __ldg((const float4 *)P.Xp + (size_t)bi * 2);
            P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, best));
            P.fxyz[pos] = nn.x; P.fxyz[(size_t)m + pos] = nn.y; P.fxyz[(size_t)2 * m + pos] = nn.z;
            P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
            icp_dist_id di; di.dist = best; di.id = bi;
            P.NNID[pos] = di;
            P.qperm[pos] = qidx(lq);
            if (walked) P.nn_o[qidx(lq)] = bi;           // seed of the next iteration's pruned walk
            if (settle)
            {
                // every other point of the list: computed distance >= sec, true sqrt(D) >= sqrt(sec) * (1 - 1e-6)
                const bool usable = (len > 0u) && (best < CUDART_INF_F) && (sec > 1e-30f);
                P.nn_o[qidx(lq)] = bi;
                P.nnd[qidx(lq)] = usable ? __fmul_rd(__fsqrt_rd(sec), 0.999999f) : -1.f;
            }
            e_cnt += len;
            x_cnt += len;
        }
    }
    PROF_STAMP(P, 2, 5, (unsigned long long)clock64()); PROF_STAMP(P, 2, 6, gtime_ns()); PROF_END_ALL(P, 2);
    if (P.evals)
    {
        unsigned long long e = e_cnt, x = x_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { e += __shfl_down_sync(FULL_MASK, e, d); x += __shfl_down_sync(FULL_MASK, x, d); }
        if (lane == 0 && e) atomicAdd(P.evals + 1, e);
        if (lane == 0 && x) atomicAdd(P.evals + 3, x);
        if (bx == 0 && threadIdx.x == 0) atomicAdd(P.evals, (unsigned long long)m * nr);
    }
}

template <bool SETTLE>
__global__ void __launch_bounds__(GROUPED_WARPS * 32, GROUPED_MINB) k_search_grouped(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ float4 smem_g4[];
    pdl_wait(); pdl_trigger();
    const PairPtrs P = table[blockIdx.y];
    search_grouped_body<SETTLE, GROUPED_WARPS>(P, cfg, blockIdx.x, smem_g4);
}

// =================================================================================================
// B' + C' (sorted flavour, Cmode 2).  B' = one CTA per pair: column prefixes of the chunk histograms (in shared
// memory), list sizes Nq, offsets Oq, and the stable sorted order itself, qperm[sorted position] = query.  C' then
// owns QG consecutive SORTED positions: its queries arrive grouped by representative (a group is cut only where a CTA
// range ends, 619 work items per pair instead of 1372 with 1024 consecutive unsorted queries per CTA), there is no
// grouping pass, no shared-memory atomics, and every output row (W, f.xyz, m.xyz, NN_ID) is written coalesced.
// The results are those of k_search<L> / k_search_grouped bit for bit (same evaluation, same ordered merge).
// =================================================================================================
#define COLSORT_THREADS 1024
// shared memory of B': Hs[nbA][nr] | cnt[nr] | sOq[nr] | part[4][nr] (partial column sums of the 4-thread scan)
#define COLSORT_PART_OFF (nr)
static size_t colsort_smem_bytes(const FusedCfg &cfg) { return ((size_t)cfg.nbA * cfg.nr + 6u * cfg.nr) * 4u + 16u; }

// REC (span flavour, Cmode 3): the scatter also moves a complete per-query record into sorted order -- the transformed
// query (32 B) and {motion-adjusted lower bound, last nearest neighbour, original index, representative} (16 B) -- so that
// kernel C'' reads everything it needs coalesced instead of gathering M, q_rep, nnd and nn_o through qperm (five dependent
// sector-granular gathers per query in k_search_sorted).  The bound update lb <- lb - delta of DESIGN 4.5 moves here with it.
template <bool MULTI, bool REC>       // MULTI: several CTAs per pair + 4 threads per column (latency mode); else one CTA per pair (batch mode)
__global__ void __launch_bounds__(COLSORT_THREADS, MULTI ? 1 : 2) k_colscan_sort(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ __align__(16) uint32_t smem_cs[];
    __shared__ uint32_t warp_tot[32];
    const PairPtrs P = table[blockIdx.y];
    const uint32_t done = __ldcg(&P.state->done);
    const uint32_t nr = cfg.nr, nb = cfg.nbA, m = cfg.m, QB = cfg.QB;
    uint32_t *Hs = smem_cs, *cnt = Hs + (size_t)nb * nr, *sOq = cnt + nr;
    const uint32_t tid = threadIdx.x;
    if (!MULTI) { for (uint32_t i = tid; i < nb * nr; i += COLSORT_THREADS) Hs[i] = __ldcg(P.H + i); }
    else
    {
        // the chunk histograms (nr % 4 == 0, 256-byte aligned): 16-byte loads, 2 in flight per thread (32-register budget)
        const uint4 *src = reinterpret_cast<const uint4 *>(P.H);
        uint4 *dst = reinterpret_cast<uint4 *>(Hs);
        const uint32_t n4 = nb * nr / 4u;
        for (uint32_t i0 = 0; i0 < n4; i0 += 2u * COLSORT_THREADS)
        {
            const uint32_t ia = i0 + tid, ib = ia + COLSORT_THREADS;
            uint4 va = make_uint4(0, 0, 0, 0), vb = va;
            if (ia < n4) va = __ldcg(src + ia);
            if (ib < n4) vb = __ldcg(src + ib);
            if (ia < n4) dst[ia] = va;
            if (ib < n4) dst[ib] = vb;
        }
    }
    __syncthreads();
    if (done) return;
    // column prefixes.  Few columns (latency mode: 147 chunk rows x 256 columns): SUB = 4 threads share a column, each
    // scans a quarter of the rows, then the quarters are chained through shared memory.
    const uint32_t SUB = (MULTI && nr * 4u <= COLSORT_THREADS && nb >= 32u) ? 4u : 1u;
    const uint32_t rows_per = (nb + SUB - 1u) / SUB;
    uint32_t *part = sOq;                                        // [SUB][nr] partial sums (sOq is not live yet; SUB * nr <= 1024 <= ... see smem size)
    for (uint32_t k = tid; k < nr * SUB; k += COLSORT_THREADS)
    {
        const uint32_t r = k % nr, sub = k / nr;
        const uint32_t row0 = sub * rows_per, row1 = min(nb, row0 + rows_per);
        uint32_t run = 0;
        for (uint32_t row = row0; row < row1; ++row) { const uint32_t t = Hs[row * nr + r]; Hs[row * nr + r] = run; run += t; }
        if (SUB == 1u) { cnt[r] = run; if (blockIdx.x == 0) P.Nq[r] = run; }
        else part[COLSORT_PART_OFF + sub * nr + r] = run;
    }
    if (MULTI && SUB > 1u)
    {
        __syncthreads();
        for (uint32_t k = tid; k < nr * SUB; k += COLSORT_THREADS)
        {
            const uint32_t r = k % nr, sub = k / nr;
            uint32_t base = 0, tot = 0;
            for (uint32_t s2 = 0; s2 < SUB; ++s2) { const uint32_t v = part[COLSORT_PART_OFF + s2 * nr + r]; if (s2 < sub) base += v; tot += v; }
            const uint32_t row0 = sub * rows_per, row1 = min(nb, row0 + rows_per);
            if (base) for (uint32_t row = row0; row < row1; ++row) Hs[row * nr + r] += base;
            if (sub == 0) { cnt[r] = tot; if (blockIdx.x == 0) P.Nq[r] = tot; }
        }
    }
    __syncthreads();
    cta_exscan_smem(cnt, nr, sOq, warp_tot);
    if (blockIdx.x == 0) for (uint32_t r = tid; r < nr; r += COLSORT_THREADS) P.Oq[r] = sOq[r];
    // every CTA scatters its slice of the queries
    const uint32_t per_cta = MULTI ? (m + gridDim.x - 1u) / gridDim.x : m;
    const uint32_t i0 = MULTI ? blockIdx.x * per_cta : 0u, i1 = MULTI ? min(m, i0 + per_cta) : m;
    if (!REC)
    {
#pragma unroll 4
        for (uint32_t i = i0 + tid; i < i1; i += COLSORT_THREADS)
        {
            const uint32_t r = __ldcg(P.q_rep + i);
            const uint32_t pos = sOq[r] + Hs[(i / QB) * nr + r] + __ldcg(P.lrank + i);
            P.QR[pos] = make_uint2(i, r);        // kernel C' reads both coalesced (and writes qperm[pos] = i for the observers)
        }
        return;
    }
    const float4 tq = __ldg((const float4 *)P.T), tt = __ldg((const float4 *)P.T + 1);
    // pose of the previous iteration (kept by kernel D) and whether the recorded bounds describe the current moving set
    const float4 pq = __ldcg((const float4 *)(P.wconst + 4)), pt = __ldcg((const float4 *)(P.wconst + 4) + 1);
    const bool settle = cfg.settle != 0;
    const bool bounds_ok = settle && __ldcg(P.wconst + 13) != 0u;
    const float fg = cfg.fg;
    float4 *Qs = reinterpret_cast<float4 *>(P.Qs);
#pragma unroll 2
    for (uint32_t i = i0 + tid; i < i1; i += COLSORT_THREADS)
    {
        const uint32_t r = __ldcg(P.q_rep + i);
        const uint32_t pos = sOq[r] + Hs[(i / QB) * nr + r] + __ldcg(P.lrank + i);
        pt8 q = ld_pt8(P.M, i);
        const float4 mlo = q.lo;
        q.lo = transform_q_xyz(q.lo, tq, tt);
        float lbn = -1.f;
        uint32_t nno = 0u;
        if (settle)
        {
            // lb <- lb - delta: delta = distance the query moved in the metric space since the bound was recorded,
            // every operation rounded against the bound
            const float lbv = __ldcg(P.nnd + i);
            nno = __ldcg(P.nn_o + i);
            if (bounds_ok && lbv > 0.f)
            {
                const float4 qp = transform_q_xyz(mlo, pq, pt);
                const float dx = fmaxf(fabsf(__fsub_ru(q.lo.x, qp.x)), fabsf(__fsub_rd(q.lo.x, qp.x)));
                const float dy = fmaxf(fabsf(__fsub_ru(q.lo.y, qp.y)), fabsf(__fsub_rd(q.lo.y, qp.y)));
                const float dz = fmaxf(fabsf(__fsub_ru(q.lo.z, qp.z)), fabsf(__fsub_rd(q.lo.z, qp.z)));
                const float s2 = __fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz));
                lbn = __fsub_rd(lbv, __fsqrt_ru(__fmul_ru(fg, s2)));
            }
        }
        P.qperm[pos] = i;
        Qs[(size_t)pos * 2] = q.lo; Qs[(size_t)pos * 2 + 1] = q.hi;
        P.Rs[pos] = make_uint4(__float_as_uint(lbn), nno, i, r);
    }
}

#ifndef SORTED_WARPS
#define SORTED_WARPS 16
#endif
#ifndef SORTED_MINB
#define SORTED_MINB 2
#endif
#ifndef SORTED_BIG_FIRST
#define SORTED_BIG_FIRST 1
#endif
#ifndef SORTED_BIG_EVALS
#define SORTED_BIG_EVALS 32u      // evaluations per lane from which a work item of C' counts as long (scheduled first)
#endif
// (Built, parity-tested, measured and removed in round 2 -- commit 47f1073, profiles/r02_ab_pass1_async_ld256.md: pass 1 with every
// dependent level of gathers batched / staged by cp.async, and the next work item's first tile prefetched.)
// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only): the gather lands in shared memory without passing through registers
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// Work items of the sorted kernel C': a group of n queries that share a list of `len` points is cut into items of <= 32
// queries; an item pads its queries to a power of two w and spreads the list over 32 / w lane phases, so a 21-query item
// wastes a third of its lanes.  Instead of ceil(n / 32) fixed slices the tail is decomposed: 21 -> 16 + 5 (padded to 8),
// 24 -> 16 + 8, ... whenever the padding would cost more evaluations than one more item's fixed overhead (cfg.item_ovh evaluations
// per lane).  Same results whatever the cut is (every query's scan is the ordered merge over its phases).
__device__ __forceinline__ uint32_t sorted_item_take(uint32_t n, uint32_t len, uint32_t ovh)
{
    if (n >= 32u) return 32u;
    const uint32_t p = 1u << (31u - (uint32_t)__clz(n));         // largest power of two <= n
    if (p == n) return n;
    const uint32_t waste = (2u * p - n) * len;                   // evaluations of the padded lanes
    return (waste <= 32u * ovh) ? n : p;
}
__device__ __forceinline__ uint32_t sorted_item_count(uint32_t n, uint32_t len, uint32_t ovh)
{
    uint32_t c = 0;
    while (n) { n -= sorted_item_take(n, len, ovh); ++c; }
    return c;
}

struct SortedSmem
{
    float4 *qlo, *qhi;          // [QG] transformed queries in sorted order
    float4 *tile;               // [SORTED_WARPS][64]
    uint32_t *sOq, *sNq, *sO, *sN, *nsl, *ibase;     // [nr] each
    uint32_t *cnt, *offC;       // [nr] each: queries of the CTA still to be searched per representative (settle flavour)
    uint32_t *items;            // [nr + QG/QI + 1]
    uint32_t *sidx, *rs;        // [QG] each (settle flavour): group slot -> local query; local query -> representative | slot << 16
};
__host__ __device__ static inline size_t sorted_carve(SortedSmem *g, void *base, uint32_t nr, uint32_t QG, uint32_t QI)
{
    char *p = (char *)base;
    size_t off = 0;
    if (g) g->qlo = (float4 *)(p + off); off += (size_t)QG * 16;
    if (g) g->qhi = (float4 *)(p + off); off += (size_t)QG * 16;
    if (g) g->tile = (float4 *)(p + off); off += (size_t)SORTED_WARPS * 64 * 16;
    uint32_t **arr[8] = { g ? &g->sOq : nullptr, g ? &g->sNq : nullptr, g ? &g->sO : nullptr, g ? &g->sN : nullptr,
                          g ? &g->nsl : nullptr, g ? &g->ibase : nullptr, g ? &g->cnt : nullptr, g ? &g->offC : nullptr };
    for (int i = 0; i < 8; ++i) { if (g) *arr[i] = (uint32_t *)(p + off); off += (size_t)nr * 4; }
    if (g) g->items = (uint32_t *)(p + off); off += ((size_t)(5u * (nr < QG ? nr : QG) + QG / 32u + 1u) * 4 + 15u) & ~(size_t)15u;   // sidx stays 16-byte aligned (cp.async destination)
    if (g) g->sidx = (uint32_t *)(p + off); off += (size_t)QG * 4;
    if (g) g->rs = (uint32_t *)(p + off); off += (size_t)QG * 4;
    return off + 16;
}

// scan flavours that also track the runner-up: sec = second smallest evaluated distance (the smallest of the others;
// equal to best for duplicates).  max/min ignore NaN operands, so a NaN distance pushes sec down to best => no settling.
template <bool FAST>
__device__ __forceinline__ void scan_tile_sec(const float4 *tlo, const float4 *thi, uint32_t tl, uint32_t ph, uint32_t Pn, uint32_t kbase,
                                              const pt8 &q, float fg, float fp, float &best, uint32_t &bi, float &sec)
{
#pragma unroll 4
    for (uint32_t k = ph; k < tl; k += Pn)
    {
        const float4 xlo = tlo[k], xhi = thi[k];
        const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        sec = fminf(sec, fmaxf(d, best));
        if (d < best) { best = d; bi = kbase + k; }
    }
}
template <bool FAST>
__device__ __forceinline__ void scan_tile_full_sec(const float4 *tlo, const float4 *thi, uint32_t kbase,
                                                   const pt8 &q, float fg, float fp, float &best, uint32_t &bi, float &sec)
{
    uint32_t bk = 0xFFFFFFFFu;
ICP_UNROLL(SCAN_FULL_UNROLL)
    for (uint32_t k = 0; k < 32u; ++k)
    {
        const float4 xlo = tlo[k], xhi = thi[k];
        const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        sec = fminf(sec, fmaxf(d, best));
        if (d < best) { best = d; bk = k; }
    }
    if (bk != 0xFFFFFFFFu) bi = kbase + bk;
}

template <int CL, int T>
__device__ __forceinline__ void reduce_solve_body(const PairPtrs &P, const FusedCfg &cfg, cudaGraphConditionalHandle handle, int use_handle,
                                  float *smem_d, const uint32_t rank, const int wide_phase = 0);

// FUSE_D: the last CTA of the pair to finish its list scans runs kernel D's body (reductions + solve + pose update).
// (Measured and dropped, round 2: launching the pair's 8 CTAs as one cluster that runs kernel D together over distributed
// shared memory -- 3x shorter per pair, but it spends 8 CTAs x 18 us instead of 1 CTA x 60 us of SM time and holds every CTA at
// the cluster barrier until the pair's slowest one is done: 8 881 against 10 093 pairs/s.)
template <bool FUSE_D>
__global__ void __launch_bounds__(SORTED_WARPS * 32, SORTED_MINB) k_search_sorted(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ float4 smem_s4[];
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_ctr;
    __shared__ unsigned long long s_lmax;                       // phase clocks only: longest item loop among the CTA's warps
    const PairPtrs P = table[blockIdx.y];
    if (__ldcg(&P.state->done)) return;
    const long long c_t0 = clock64();
    const uint32_t nr = cfg.nr, m = cfg.m, QI = cfg.QI, QG = cfg.QG;
    const bool settle = cfg.settle != 0;
    SortedSmem G;
    sorted_carve(&G, smem_s4, nr, QG, QI);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t p0 = blockIdx.x * QG, nq_cta = min(QG, m - p0), p1 = p0 + nq_cta;
    for (uint32_t r = tid; r < nr; r += blockDim.x)
    {
        const uint32_t oq = __ldcg(P.Oq + r), nq = __ldcg(P.Nq + r);
        G.sOq[r] = oq; G.sNq[r] = nq;
        G.sO[r] = __ldg(P.O + r);
        G.sN[r] = __ldg(P.N + r);
        const uint32_t lo = max(oq, p0), hi = min(oq + nq, p1);
        G.nsl[r] = hi > lo ? sorted_item_count(hi - lo, G.sN[r], cfg.item_ovh) : 0u;
        G.cnt[r] = 0u;
    }
    if (tid == 0) { s_ctr = 0; s_lmax = 0ull; }
    const float4 tq = __ldg((const float4 *)P.T), tt = __ldg((const float4 *)P.T + 1);
    // pose of the previous iteration (kernel D keeps it): the settle test measures how far every query moved since then
    const float4 pq = __ldcg((const float4 *)(P.wconst + 4)), pt = __ldcg((const float4 *)(P.wconst + 4) + 1);
    const float w_lo = __ldg(P.reps + 3), w_hi = __ldg(P.reps + 7);
    const float fg = cfg.fg, fp = cfg.fp;
    bool fast = __ldcg(P.wconst) != 0u;
    const bool bounds_ok = settle && __ldcg(P.wconst + 13) != 0u;
    unsigned long long e_cnt = 0, x_cnt = 0;
    long long c_ta = 0, c_tb = 0;
    {
    // pass 1: {original index, representative} of the CTA's sorted positions arrive coalesced (B' wrote them); per query the
    // gathers are the point M[i] and, with the temporal pruning, nnd[i], nn_o[i] and then X_p[nn_o[i]].  The addresses of all
    // the thread's queries are issued up front (prefetches carry no destination register), the loop below then runs out of
    // L1 / L2.  (Measured and dropped, round 2: one 48-byte state record per query holding the bound, the position AND the
    // coordinates of x* -- one gather instead of three, two dependent latencies instead of three -- made pass 1 10 % slower:
    // the records are 786 KB of cold DRAM per pair and iteration, while X_p[nn_o] mostly hits the L2 lines the list scans use.)
    constexpr uint32_t PF = 4;
    {
        uint32_t pf_n[PF];
#pragma unroll
        for (uint32_t j = 0; j < PF; ++j)
        {
            const uint32_t l = tid + j * blockDim.x;
            pf_n[j] = 0xFFFFFFFFu;
            if (l < nq_cta)
            {
                const uint32_t i = __ldcg(&P.QR[p0 + l].x);
                prefetch_l1(reinterpret_cast<const float4 *>(P.M) + (size_t)i * 2);
                if (settle) pf_n[j] = __ldcg(&P.nn2[i].y);
            }
        }
        if (settle)
        {
#pragma unroll
            for (uint32_t j = 0; j < PF; ++j)
                if (pf_n[j] < m) prefetch_l1(reinterpret_cast<const float4 *>(P.Xp) + (size_t)pf_n[j] * 2);
        }
    }
    c_ta = clock64();
    if (settle) __syncthreads();                 // sO / sN / cnt are used by pass 1
    c_tb = clock64();
    for (uint32_t l = tid; l < nq_cta; l += blockDim.x)
    {
        const uint2 ir = __ldcg(P.QR + p0 + l);
        const uint32_t i = ir.x, r = ir.y;
        P.qperm[p0 + l] = i;
        pt8 q = ld_pt8(P.M, i);
        const float4 mlo = q.lo;
        q.lo = transform_q_xyz(q.lo, tq, tt);
        fast = fast && (q.lo.w == w_lo) && (q.hi.w == w_hi);
        G.qlo[l] = q.lo; G.qhi[l] = q.hi;
        if (settle)
        {
            // Exact temporal pruning of stage 2.  lb = proven lower bound of sqrt(D) between this query and every
            // point of its list other than last iteration's nearest neighbour x*; the query moved by at most delta in the
            // metric space since then, so every other point is still farther than lb - delta.  If even the COMPUTED distance
            // of any other point (>= true * (1 - 1e-6) - tiny) must exceed the computed distance to x*, the sequential scan
            // would return x* again: evaluate that one distance with the reference arithmetic and skip the scan.
            const uint2 bn = __ldcg(P.nn2 + i);
            const float lbv = __uint_as_float(bn.x);
            const uint32_t nno = bn.y;
            const uint32_t o = G.sO[r], len = G.sN[r];
            bool settled = false;
            if (bounds_ok && lbv > 0.f && (nno - o) < len)       // same representative as when x* was found (lists are disjoint)
            {
                const float4 qp = transform_q_xyz(mlo, pq, pt);
                const float dx = fmaxf(fabsf(__fsub_ru(q.lo.x, qp.x)), fabsf(__fsub_rd(q.lo.x, qp.x)));
                const float dy = fmaxf(fabsf(__fsub_ru(q.lo.y, qp.y)), fabsf(__fsub_rd(q.lo.y, qp.y)));
                const float dz = fmaxf(fabsf(__fsub_ru(q.lo.z, qp.z)), fabsf(__fsub_rd(q.lo.z, qp.z)));
                const float s2 = __fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz));
                const float delta = __fsqrt_ru(__fmul_ru(fg, s2));
                const float lbn = __fsub_rd(lbv, delta);
                const pt8 x = ld_pt8(P.Xp, nno);
                const float d = dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);     // == dist6 bit for bit whenever dist6 applies
                if (lbn > 0.f && __fmul_rd(__fmul_rd(lbn, lbn), 0.99999f) > __fadd_ru(d, 1e-30f))
                {
                    settled = true;
                    const uint32_t pos = p0 + l;
                    P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, d));
                    P.fxyz[pos] = x.lo.x; P.fxyz[(size_t)m + pos] = x.lo.y; P.fxyz[(size_t)2 * m + pos] = x.lo.z;
                    P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
                    icp_dist_id di; di.dist = d; di.id = nno;
                    P.NNID[pos] = di;
                    P.nn2[i].x = __float_as_uint(lbn);
                    e_cnt += len;
                    x_cnt += 1u;
                }
            }
            if (settled) G.rs[l] = 0xFFFFFFFFu;
            else
            {
                const uint32_t slot = atomicAdd(&G.cnt[r], 1u);
                G.rs[l] = r | (slot << 16);
            }
        }
    }
    }
    fast = __syncthreads_and(fast) != 0;
    const long long c_t1 = clock64();
    // Work items: ONE pass builds both prefix sums (slots of the unsettled queries per representative -> offC, items per
    // representative -> ibase) and the items themselves -- thread t owns `per` consecutive representatives, the warps chain
    // through two shuffle scans and one table of warp totals (2 barriers instead of 9 for the two separate scans + loops;
    // the phase clocks of round 2 put this part at 14 % of the kernel).
    __shared__ uint32_t wt_c[32], wt_i[32], wt_b[32];
    uint32_t nitems;
    {
        // Longest items first: the warps take the items in array order, so whatever is taken last sticks out of the CTA (measured:
        // the last warp finished 6.1 K cycles after the mean, 38.6 K against 32.4 K).  Items whose scan costs >= SORTED_BIG_EVALS
        // evaluations per lane go to the front of the array, the others behind them (a third prefix sum; any order is correct).
        auto item_is_big = [&](uint32_t c, uint32_t len) -> bool {
            const uint32_t w = c > 1u ? 1u << (32u - (uint32_t)__clz(c - 1u)) : 1u;         // queries padded to a power of two
            return len * w >= 32u * SORTED_BIG_EVALS;                                       // = list points per lane
        };
        const uint32_t nthreads = blockDim.x, nwarps = nthreads >> 5;
        const uint32_t per = (nr + nthreads - 1u) / nthreads;
        const uint32_t b0 = tid * per;
        uint32_t sum_c = 0, sum_i = 0, sum_b = 0;
        for (uint32_t j = 0; j < per; ++j)
        {
            const uint32_t r = b0 + j;
            if (r >= nr) break;
            uint32_t n;
            if (settle) n = G.cnt[r];
            else { const uint32_t lo = max(G.sOq[r], p0), hi = min(G.sOq[r] + G.sNq[r], p1); n = hi > lo ? hi - lo : 0u; }
            sum_c += n;
            const uint32_t len = G.sN[r];
            while (n)
            {
                const uint32_t c = sorted_item_take(n, len, cfg.item_ovh);
                ++sum_i;
                if (SORTED_BIG_FIRST && item_is_big(c, len)) ++sum_b;
                n -= c;
            }
        }
        uint32_t inc_c = sum_c, inc_i = sum_i, inc_b = sum_b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t vc = __shfl_up_sync(FULL_MASK, inc_c, d), vi = __shfl_up_sync(FULL_MASK, inc_i, d), vb = __shfl_up_sync(FULL_MASK, inc_b, d);
            if (lane >= (uint32_t)d) { inc_c += vc; inc_i += vi; inc_b += vb; }
        }
        if (lane == 31) { wt_c[warp] = inc_c; wt_i[warp] = inc_i; wt_b[warp] = inc_b; }
        __syncthreads();
        const uint32_t nact = min(nwarps, ((nr + per - 1u) / per + 31u) >> 5);       // warps that hold representatives
        uint32_t base_c = 0, base_i = 0, base_b = 0, tot_i = 0, tot_b = 0;
        for (uint32_t w2 = 0; w2 < nact; ++w2)
        {
            const uint32_t tc = wt_c[w2], ti = wt_i[w2], tb = wt_b[w2];
            if (w2 < warp) { base_c += tc; base_i += ti; base_b += tb; }
            tot_i += ti; tot_b += tb;
        }
        nitems = tot_i;
        uint32_t run_c = base_c + inc_c - sum_c;
        uint32_t run_b = base_b + inc_b - sum_b;                                             // next big item of this thread
        uint32_t run_s = tot_b + (base_i - base_b) + (inc_i - inc_b) - (sum_i - sum_b);      // next small item
        // item = representative | first slot of the group << 12 | queries << 24   (nr <= 4096, slots < 4096, queries <= 32)
        for (uint32_t j = 0; j < per; ++j)
        {
            const uint32_t r = b0 + j;
            if (r >= nr) break;
            uint32_t n;
            if (settle) n = G.cnt[r];
            else { const uint32_t lo = max(G.sOq[r], p0), hi = min(G.sOq[r] + G.sNq[r], p1); n = hi > lo ? hi - lo : 0u; }
            G.offC[r] = run_c;
            run_c += n;
            const uint32_t len = G.sN[r];
            uint32_t slot = 0;
            while (n)
            {
                const uint32_t c = sorted_item_take(n, len, cfg.item_ovh);
                const uint32_t item = r | (slot << 12) | (c << 24);
                if (SORTED_BIG_FIRST && item_is_big(c, len)) G.items[run_b++] = item; else G.items[run_s++] = item;
                slot += c; n -= c;
            }
        }
        __syncthreads();
    }
    if (settle)
        for (uint32_t l = tid; l < nq_cta; l += blockDim.x)
        {
            const uint32_t v = G.rs[l];
            if (v != 0xFFFFFFFFu) G.sidx[G.offC[v & 0xFFFFu] + (v >> 16)] = l;
        }
    __syncthreads();
    const long long c_t2 = clock64();

    float4 *tlo = G.tile + warp * 64u, *thi = tlo + 32;
    while (true)
    {
        uint32_t it = 0;
        if (lane == 0) it = atomicAdd(&s_ctr, 1u);
        it = __shfl_sync(FULL_MASK, it, 0);
        if (it >= nitems) break;
        const uint32_t item = G.items[it];
        const uint32_t r = item & 0xFFFu, slot0 = (item >> 12) & 0xFFFu, nq = item >> 24;
        uint32_t l0 = 0;
        if (!settle) l0 = max(G.sOq[r], p0) - p0 + slot0;        // first local query of the item (the group's part inside this CTA)
        const uint32_t lw = nq > 1u ? 32u - (uint32_t)__clz(nq - 1u) : 0u;
        const uint32_t w = 1u << lw;                             // queries (padded to a power of two) ...
        const uint32_t Pn = 32u >> lw;                           // ... x list phases
        const uint32_t ql = lane & (w - 1u), ph = lane >> lw;
        const bool valid = ql < nq;
        const uint32_t lq = settle ? G.sidx[G.offC[r] + slot0 + (valid ? ql : 0u)] : l0 + (valid ? ql : 0u);
        pt8 q; q.lo = G.qlo[lq]; q.hi = G.qhi[lq];
        const uint32_t o = G.sO[r], len = G.sN[r];
        float best = CUDART_INF_F, sec = CUDART_INF_F;
        uint32_t bi = o;
        pt8 nx;
        if (lane < len) nx = ld_pt8(P.Xp, o + lane);
        for (uint32_t t0 = 0; t0 < len; t0 += 32u)
        {
            const uint32_t tl = min(32u, len - t0);
            __syncwarp();
            if (lane < tl) { tlo[lane] = nx.lo; thi[lane] = nx.hi; }
            __syncwarp();
            if (t0 + 32u + lane < len) nx = ld_pt8(P.Xp, o + t0 + 32u + lane);
            if (settle)
            {
                if (Pn == 1u && tl == 32u)
                {
                    if (fast) scan_tile_full_sec<true>(tlo, thi, o + t0, q, fg, fp, best, bi, sec);
                    else scan_tile_full_sec<false>(tlo, thi, o + t0, q, fg, fp, best, bi, sec);
                }
                else if (fast) scan_tile_sec<true>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi, sec);
                else scan_tile_sec<false>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi, sec);
            }
            else if (Pn == 1u && tl == 32u)
            {
                if (fast) scan_tile_full<true>(tlo, thi, o + t0, q, fg, fp, best, bi);
                else scan_tile_full<false>(tlo, thi, o + t0, q, fg, fp, best, bi);
            }
            else if (fast) scan_tile<true>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi);
            else scan_tile<false>(tlo, thi, tl, ph, Pn, o + t0, q, fg, fp, best, bi);
        }
        for (uint32_t off = w; off < 32u; off <<= 1)
        {
            const float od = __shfl_xor_sync(FULL_MASK, best, off);
            const uint32_t oi = __shfl_xor_sync(FULL_MASK, bi, off);
            const float os = __shfl_xor_sync(FULL_MASK, sec, off);
            sec = fminf(fminf(sec, os), fmaxf(best, od));        // runner-up of the union of the two phases
            if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
        }
        if (valid && ph == 0)
        {
            if (best == CUDART_INF_F) bi = o;           // nothing compared less than +inf: the sequential scan keeps the list head
            if (len == 0) bi = o ? o - 1u : 0u;
            if (bi >= m) bi = m - 1u;
            const uint32_t pos = p0 + lq;
            const float4 nn = __ldg((const float4 *)P.Xp + (size_t)bi * 2);
            P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, best));
            P.fxyz[pos] = nn.x; P.fxyz[(size_t)m + pos] = nn.y; P.fxyz[(size_t)2 * m + pos] = nn.z;
            P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
            icp_dist_id di; di.dist = best; di.id = bi;
            P.NNID[pos] = di;
            if (settle)
            {
                // every other point of the list: computed distance >= sec, true sqrt(D) >= sqrt(sec) * (1 - 1e-6)
                const uint32_t i = __ldcg(&P.QR[pos].x);
                const bool usable = (len > 0u) && (best < CUDART_INF_F) && (sec > 1e-30f);
                P.nn2[i] = make_uint2(__float_as_uint(usable ? __fmul_rd(__fsqrt_rd(sec), 0.999999f) : -1.f), bi);
            }
            e_cnt += len;
            x_cnt += len;
        }
    }
    if (P.prof && lane == 0)
    {
        // phase clocks of this launch, summed over the pair's CTAs: [32] set-up + pass 1 (thread 0), [33] item build,
        // [34] item loop of every warp (warp-cycles: divide by the warps), [35] item loop until the CTA's LAST warp is done
        const long long c_t3 = clock64();
        if (tid == 0) { atomicAdd(P.prof + 38, (unsigned long long)(c_ta - c_t0)); atomicAdd(P.prof + 39, (unsigned long long)(c_tb - c_ta)); atomicAdd(P.prof + 32, (unsigned long long)(c_t1 - c_t0)); atomicAdd(P.prof + 33, (unsigned long long)(c_t2 - c_t1)); atomicAdd(P.prof + 36, 1ull); }
        atomicAdd(P.prof + 34, (unsigned long long)(c_t3 - c_t2));
        atomicMax(P.prof + 37, (unsigned long long)(c_t3 - c_t0));
        atomicMax(&s_lmax, (unsigned long long)(c_t3 - c_t2));
    }
    if (P.evals)
    {
        unsigned long long e = e_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(FULL_MASK, e, d);
        unsigned long long x = x_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(FULL_MASK, x, d);
        if (lane == 0 && e) atomicAdd(P.evals + 1, e);
        if (lane == 0 && x) atomicAdd(P.evals + 3, x);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.evals, (unsigned long long)m * nr);
    }
    if (FUSE_D)
    {
        // every thread publishes its outputs, then one arrival per CTA; the last CTA of the pair continues with kernel D
        __shared__ uint32_t s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0 && P.prof) atomicAdd(P.prof + 35, s_lmax);
        if (tid == 0)
        {
            const uint32_t prev = atomicAdd(P.wconst + 2, 1u);
            s_last = (prev + 1u == gridDim.x) ? 1u : 0u;
            if (s_last) P.wconst[2] = 0u;                    // re-armed for the next iteration (nobody else touches it until then)
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        reduce_solve_body<1, SORTED_WARPS * 32>(P, cfg, 0, 0, reinterpret_cast<float *>(smem_s4), 0u);
    }
}

// =================================================================================================
// C'' (span flavour, Cmode 3): k_search_sorted re-cut around two facts measured on B200 (profiles/r02_*): (1) a third of
// C' was pass 1 -- five dependent, sector-granular gathers per query (qperm -> M, q_rep, nnd, nn_o -> X_p[nn_o]); (2) its
// item loop paid ~300 instructions and two L2 latencies per work item to stream the list through a per-warp tile.
//   * B'' (k_colscan_sort<., true>) already scattered a complete record per query into sorted order, so pass 1 is two
//     coalesced loads per query;
//   * a CTA owns QG consecutive sorted positions = a run of consecutive representatives, whose lists are ONE contiguous
//     span of X_p: it is staged into shared memory once, by bulk-async copies (cp.async.bulk + mbarrier: the TMA unit
//     moves the bytes while the CTA loads its query records), and every list scan, the settle test's x* and the epilogue's
//     matched point read it by broadcast LDS.128 -- no tiles, no warp barriers, no global latency inside the item loop;
//   * a span larger than the window is worked off in rounds of whole lists; a single list larger than the window (degenerate
//     clouds: thousands of identical points) is scanned straight from global memory (same arithmetic, L1-served).
// Results are those of k_search_sorted bit for bit: same distance arithmetic, same ordered merges, same bounds.
// =================================================================================================
#define SPAN_THREADS 1024
#define SPAN_CHUNK_BYTES 16384u
struct SpanSmem
{
    float4 *qlo, *qhi;          // [QG] transformed queries in sorted order
    float4 *span;               // [2 * span_pts] window of X_p, AoS exactly as in global memory
    uint32_t *sOq, *sNq, *sO, *sN, *nsl, *ibase, *cnt, *offC;     // [nr] each
    uint32_t *items;            // [nr + QG/QI + 1]
    uint32_t *sidx, *rs, *qi;   // [QG] each: group slot -> local query; local query -> representative | slot << 16; original index
};
__host__ __device__ static inline size_t span_carve(SpanSmem *g, void *base, uint32_t nr, uint32_t QG, uint32_t QI, uint32_t span_pts)
{
    char *p = (char *)base;
    size_t off = 0;
    if (g) g->span = (float4 *)(p + off); off += (size_t)span_pts * 32;        // first: 128-byte aligned destination of the bulk copies
    if (g) g->qlo = (float4 *)(p + off); off += (size_t)QG * 16;
    if (g) g->qhi = (float4 *)(p + off); off += (size_t)QG * 16;
    uint32_t **arr[8] = { g ? &g->sOq : nullptr, g ? &g->sNq : nullptr, g ? &g->sO : nullptr, g ? &g->sN : nullptr,
                          g ? &g->nsl : nullptr, g ? &g->ibase : nullptr, g ? &g->cnt : nullptr, g ? &g->offC : nullptr };
    for (int i = 0; i < 8; ++i) { if (g) *arr[i] = (uint32_t *)(p + off); off += (size_t)nr * 4; }
    if (g) g->items = (uint32_t *)(p + off); off += (size_t)(nr + QG / QI + 1) * 4;
    if (g) g->sidx = (uint32_t *)(p + off); off += (size_t)QG * 4;
    if (g) g->rs = (uint32_t *)(p + off); off += (size_t)QG * 4;
    if (g) g->qi = (uint32_t *)(p + off); off += (size_t)QG * 4;
    return off + 16;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tSPAN_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra SPAN_DONE;\n\tbra SPAN_WAIT;\n\tSPAN_DONE:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// one list (len points at L, AoS) against the lane's query; lane phase ph of PN scans points ph, ph + PN, ...
// SEC: also track the runner-up (see scan_tile_sec).  bk = winning position relative to the list start, or 0xFFFFFFFF.
template <bool FAST, bool SEC, int PN>
__device__ __forceinline__ void scan_span(const float4 *__restrict__ L, uint32_t len, uint32_t ph, const pt8 &q, float fg, float fp,
                                          float &best, uint32_t &bk, float &sec)
{
    constexpr int U = (PN == 1) ? 8 : 4;
    uint32_t k = ph;
    for (; k + (uint32_t)((U - 1) * PN) < len; k += (uint32_t)(U * PN))
    {
        const float4 *Lk = L + (size_t)k * 2;
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            const float4 xlo = Lk[2 * u * PN], xhi = Lk[2 * u * PN + 1];
            const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
            if (SEC) sec = fminf(sec, fmaxf(d, best));
            if (d < best) { best = d; bk = k + (uint32_t)(u * PN); }
        }
    }
    for (; k < len; k += (uint32_t)PN)
    {
        const float4 xlo = L[(size_t)k * 2], xhi = L[(size_t)k * 2 + 1];
        const float d = FAST ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        if (SEC) sec = fminf(sec, fmaxf(d, best));
        if (d < best) { best = d; bk = k; }
    }
}
template <bool FAST, bool SEC>
__device__ __forceinline__ void scan_span_any(const float4 *__restrict__ L, uint32_t len, uint32_t ph, uint32_t lw, const pt8 &q, float fg, float fp,
                                              float &best, uint32_t &bk, float &sec)
{
    switch (lw)        // lw = log2(queries of the item, padded): 32 >> lw list phases
    {
        case 5: scan_span<FAST, SEC, 1>(L, len, ph, q, fg, fp, best, bk, sec); break;
        case 4: scan_span<FAST, SEC, 2>(L, len, ph, q, fg, fp, best, bk, sec); break;
        case 3: scan_span<FAST, SEC, 4>(L, len, ph, q, fg, fp, best, bk, sec); break;
        case 2: scan_span<FAST, SEC, 8>(L, len, ph, q, fg, fp, best, bk, sec); break;
        case 1: scan_span<FAST, SEC, 16>(L, len, ph, q, fg, fp, best, bk, sec); break;
        default: scan_span<FAST, SEC, 32>(L, len, ph, q, fg, fp, best, bk, sec); break;
    }
}
// the same scan with a run-time phase count, from shared OR global memory (generic pointer): the 8-lane distance path
// (non-constant homogeneous lanes) and lists that do not fit the window (scanned where they lie, L1-served)
template <bool SEC>
__device__ __forceinline__ void scan_rt(const float4 *L, uint32_t len, uint32_t ph, uint32_t Pn, const pt8 &q, bool fast,
                                        float fg, float fp, float &best, uint32_t &bk, float &sec)
{
#pragma unroll 4
    for (uint32_t k = ph; k < len; k += Pn)
    {
        const float4 xlo = L[(size_t)k * 2], xhi = L[(size_t)k * 2 + 1];
        const float d = fast ? dist6(q.lo, q.hi, xlo, xhi, fg, fp) : dist8(q.lo, q.hi, xlo, xhi, fg, fp);
        if (SEC) sec = fminf(sec, fmaxf(d, best));
        if (d < best) { best = d; bk = k; }
    }
}

template <bool FUSE_D, int THREADS>          // 1024 threads: one CTA per SM owns 2048 sorted positions; 512 threads: two CTAs per SM own 1024 each
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_search_span(const PairPtrs *__restrict__ table, const FusedCfg cfg)
{
    extern __shared__ __align__(128) float4 smem_sp4[];
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t s_ctr;
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_win[4];            // window of the current round: first representative, one past the last, first list position, points
    const PairPtrs P = table[blockIdx.y];
    if (__ldcg(&P.state->done)) return;
    const uint32_t nr = cfg.nr, m = cfg.m, QI = cfg.QI, QG = cfg.QG, span_pts = cfg.span_pts;
    const bool settle = cfg.settle != 0;
    SpanSmem G;
    span_carve(&G, smem_sp4, nr, QG, QI, span_pts);
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t p0 = blockIdx.x * QG, nq_cta = min(QG, m - p0), p1 = p0 + nq_cta;
    for (uint32_t r = tid; r < nr; r += blockDim.x)
    {
        G.sOq[r] = __ldcg(P.Oq + r); G.sNq[r] = __ldcg(P.Nq + r);
        G.sO[r] = __ldg(P.O + r); G.sN[r] = __ldg(P.N + r);
        G.cnt[r] = 0u;
    }
    if (tid == 0) { s_ctr = 0u; mbar_init(&s_bar, 1u); }
    __syncthreads();
    // representatives of the first and the last query of the CTA: the last r with Oq[r] <= position (empty groups share
    // their successor's offset and sort before it)
    uint32_t r_next = 0, r_last = 0;         // thread 0 only
    auto stage_window = [&](uint32_t r_lo) -> uint32_t       // thread 0: choose the next run of whole lists, start its copies
    {
        uint32_t r_hi = r_lo, tot = 0u;
        while (r_hi <= r_last)
        {
            const uint32_t n = G.sN[r_hi];
            if (n > span_pts) { if (r_hi == r_lo) ++r_hi; break; }      // oversize list: a round of its own, scanned from global memory
            if (tot + n > span_pts) break;
            tot += n; ++r_hi;
        }
        const uint32_t w0 = G.sO[r_lo];
        s_win[0] = r_lo; s_win[1] = r_hi; s_win[2] = w0; s_win[3] = tot;
        if (tot > 0u)
        {
            const uint32_t bytes = tot * 32u;
            mbar_arrive_expect_tx(&s_bar, bytes);
            const char *src = reinterpret_cast<const char *>(P.Xp) + (size_t)w0 * 32u;
            char *dst = reinterpret_cast<char *>(G.span);
            for (uint32_t b = 0; b < bytes; b += SPAN_CHUNK_BYTES)
                bulk_g2s(dst + b, src + b, min(SPAN_CHUNK_BYTES, bytes - b), &s_bar);
        }
        return r_hi;
    };
    if (tid == 0)
    {
        uint32_t lo = 0, hi = nr;            // upper bound of p0 in sOq
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (G.sOq[mid] <= p0) lo = mid + 1; else hi = mid; }
        const uint32_t r_first = lo - 1u;
        lo = 0; hi = nr;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (G.sOq[mid] <= p1 - 1u) lo = mid + 1; else hi = mid; }
        r_last = lo - 1u;
        r_next = stage_window(r_first);
    }
    const float w_lo = __ldg(P.reps + 3), w_hi = __ldg(P.reps + 7);
    const float fg = cfg.fg, fp = cfg.fp;
    bool fast = __ldcg(P.wconst) != 0u;
    unsigned long long e_cnt = 0, x_cnt = 0;
    // ---- pass 1: the CTA's query records, coalesced (two queries per thread at QG = 2048); the bulk copies run meanwhile
    constexpr int QPT1 = 2;
    uint4 rec[QPT1];
    pt8 qq[QPT1];
    const float4 *Qs = reinterpret_cast<const float4 *>(P.Qs);
    for (uint32_t l0 = 0; l0 < nq_cta; l0 += QPT1 * THREADS)
    {
#pragma unroll
        for (int j = 0; j < QPT1; ++j)
        {
            const uint32_t l = l0 + (uint32_t)j * THREADS + tid;
            if (l < nq_cta)
            {
                rec[j] = __ldcg(P.Rs + p0 + l);
                qq[j].lo = __ldcg(Qs + (size_t)(p0 + l) * 2); qq[j].hi = __ldcg(Qs + (size_t)(p0 + l) * 2 + 1);
            }
        }
#pragma unroll
        for (int j = 0; j < QPT1; ++j)
        {
            const uint32_t l = l0 + (uint32_t)j * THREADS + tid;
            if (l < nq_cta)
            {
                G.qlo[l] = qq[j].lo; G.qhi[l] = qq[j].hi; G.qi[l] = rec[j].z;
                fast = fast && (qq[j].lo.w == w_lo) && (qq[j].hi.w == w_hi);
            }
        }
        if (l0 == 0u)
        {
            __syncthreads();                 // s_win of the first window is published
            if (s_win[3] > 0u) mbar_wait(&s_bar, 0u);
        }
        const uint32_t w0 = s_win[2], wn = s_win[3];
#pragma unroll
        for (int j = 0; j < QPT1; ++j)
        {
            const uint32_t l = l0 + (uint32_t)j * THREADS + tid;
            if (l >= nq_cta) continue;
            const uint32_t r = rec[j].w, i = rec[j].z, nno = rec[j].y;
            const float lbn = __uint_as_float(rec[j].x);
            const uint32_t o = G.sO[r], len = G.sN[r];
            bool settled = false;
            // exact temporal pruning of stage 2 (the bound was lowered by this iteration's motion in B''):
            // x* must still lie in the query's list (same representative as when it was found)
            if (settle && lbn > 0.f && (nno - o) < len)
            {
                pt8 x;
                if (nno - w0 < wn) { x.lo = G.span[(size_t)(nno - w0) * 2]; x.hi = G.span[(size_t)(nno - w0) * 2 + 1]; }
                else x = ld_pt8(P.Xp, nno);
                const float d = dist8(qq[j].lo, qq[j].hi, x.lo, x.hi, fg, fp);     // == dist6 bit for bit whenever dist6 applies
                if (__fmul_rd(__fmul_rd(lbn, lbn), 0.99999f) > __fadd_ru(d, 1e-30f))
                {
                    settled = true;
                    const uint32_t pos = p0 + l;
                    P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, d));
                    P.fxyz[pos] = x.lo.x; P.fxyz[(size_t)m + pos] = x.lo.y; P.fxyz[(size_t)2 * m + pos] = x.lo.z;
                    P.mxyz[pos] = qq[j].lo.x; P.mxyz[(size_t)m + pos] = qq[j].lo.y; P.mxyz[(size_t)2 * m + pos] = qq[j].lo.z;
                    icp_dist_id di; di.dist = d; di.id = nno;
                    P.NNID[pos] = di;
                    P.nnd[i] = lbn;
                    e_cnt += len;
                    x_cnt += 1u;
                }
            }
            if (settled) G.rs[l] = 0xFFFFFFFFu;
            else
            {
                const uint32_t slot = atomicAdd(&G.cnt[r], 1u);
                G.rs[l] = r | (slot << 16);
            }
        }
    }
    fast = __syncthreads_and(fast) != 0;
    for (uint32_t r = tid; r < nr; r += blockDim.x) G.nsl[r] = (G.cnt[r] + QI - 1u) / QI;
    __syncthreads();
    cta_exscan_smem(G.cnt, nr, G.offC, warp_tot);
    const uint32_t nitems = cta_exscan_smem(G.nsl, nr, G.ibase, warp_tot);
    for (uint32_t r = tid; r < nr; r += blockDim.x)
        for (uint32_t sl = 0; sl < G.nsl[r]; ++sl) G.items[G.ibase[r] + sl] = r | (sl << 16);
    for (uint32_t l = tid; l < nq_cta; l += blockDim.x)
    {
        const uint32_t v = G.rs[l];
        if (v != 0xFFFFFFFFu) G.sidx[G.offC[v & 0xFFFFu] + (v >> 16)] = l;
    }
    __syncthreads();

    // ---- rounds over windows of whole lists (one round unless the span exceeds the window)
    uint32_t it = 0;
    bool have = false;
    uint32_t parity = 0u;
    while (true)
    {
        const uint32_t w_rhi = s_win[1], w0 = s_win[2], wn = s_win[3];
        const uint32_t it_end = (w_rhi < nr) ? G.ibase[w_rhi] : nitems;       // the items are ordered by representative
        while (true)
        {
            if (!have)
            {
                if (lane == 0) it = atomicAdd(&s_ctr, 1u);
                it = __shfl_sync(FULL_MASK, it, 0);
                have = true;
            }
            if (it >= it_end) break;         // an item of a later round stays with this warp
            have = false;
            const uint32_t item = G.items[it];
            const uint32_t r = item & 0xFFFFu, sl = item >> 16;
            const uint32_t nq = min(QI, G.cnt[r] - sl * QI);
            const uint32_t lw = nq > 1u ? 32u - (uint32_t)__clz(nq - 1u) : 0u;
            const uint32_t w = 1u << lw;                             // queries (padded to a power of two) ...
            const uint32_t Pn = 32u >> lw;                           // ... x list phases
            const uint32_t ql = lane & (w - 1u), ph = lane >> lw;
            const bool valid = ql < nq;
            const uint32_t lq = G.sidx[G.offC[r] + sl * QI + (valid ? ql : 0u)];
            pt8 q; q.lo = G.qlo[lq]; q.hi = G.qhi[lq];
            const uint32_t o = G.sO[r], len = G.sN[r];
            float best = CUDART_INF_F, sec = CUDART_INF_F;
            uint32_t bk = 0xFFFFFFFFu;
            const bool in_win = (o - w0) < wn || len == 0u;          // whole lists are staged: the first point decides
            const float4 *L = in_win ? G.span + (size_t)(o - w0) * 2 : reinterpret_cast<const float4 *>(P.Xp) + (size_t)o * 2;
            if (in_win && fast)
            {
                if (settle) scan_span_any<true, true>(L, len, ph, lw, q, fg, fp, best, bk, sec);
                else scan_span_any<true, false>(L, len, ph, lw, q, fg, fp, best, bk, sec);
            }
            else if (settle) scan_rt<true>(L, len, ph, Pn, q, fast, fg, fp, best, bk, sec);
            else scan_rt<false>(L, len, ph, Pn, q, fast, fg, fp, best, bk, sec);
            uint32_t bi = (bk != 0xFFFFFFFFu) ? o + bk : o;
            for (uint32_t off = w; off < 32u; off <<= 1)
            {
                const float od = __shfl_xor_sync(FULL_MASK, best, off);
                const uint32_t oi = __shfl_xor_sync(FULL_MASK, bi, off);
                const float os = __shfl_xor_sync(FULL_MASK, sec, off);
                sec = fminf(fminf(sec, os), fmaxf(best, od));        // runner-up of the union of the two phases
                if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
            }
            if (valid && ph == 0)
            {
                if (best == CUDART_INF_F) bi = o;       // nothing compared less than +inf: the sequential scan keeps the list head
                if (len == 0) bi = o ? o - 1u : 0u;
                if (bi >= m) bi = m - 1u;
                const uint32_t pos = p0 + lq;
                float4 nn;
                if (bi - w0 < wn) nn = G.span[(size_t)(bi - w0) * 2];
                else nn = __ldg((const float4 *)P.Xp + (size_t)bi * 2);
                P.W[pos] = __fdiv_rn(100.f, __fadd_rn(100.f, best));
                P.fxyz[pos] = nn.x; P.fxyz[(size_t)m + pos] = nn.y; P.fxyz[(size_t)2 * m + pos] = nn.z;
                P.mxyz[pos] = q.lo.x; P.mxyz[(size_t)m + pos] = q.lo.y; P.mxyz[(size_t)2 * m + pos] = q.lo.z;
                icp_dist_id di; di.dist = best; di.id = bi;
                P.NNID[pos] = di;
                if (settle)
                {
                    // every other point of the list: computed distance >= sec, true sqrt(D) >= sqrt(sec) * (1 - 1e-6)
                    const uint32_t i = G.qi[lq];
                    const bool usable = (len > 0u) && (best < CUDART_INF_F) && (sec > 1e-30f);
                    P.nn_o[i] = bi;
                    P.nnd[i] = usable ? __fmul_rd(__fsqrt_rd(sec), 0.999999f) : -1.f;
                }
                e_cnt += len;
                x_cnt += len;
            }
        }
        if (it_end >= nitems) break;         // uniform: every warp saw the same window
        __syncthreads();                     // everybody is done reading the window
        if (tid == 0) r_next = stage_window(r_next);
        parity ^= (wn > 0u) ? 1u : 0u;       // the barrier completed a phase only if the last window carried bytes
        __syncthreads();
        if (s_win[3] > 0u) mbar_wait(&s_bar, parity);
    }
    if (P.evals)
    {
        unsigned long long e = e_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) e += __shfl_down_sync(FULL_MASK, e, d);
        unsigned long long x = x_cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(FULL_MASK, x, d);
        if (lane == 0 && e) atomicAdd(P.evals + 1, e);
        if (lane == 0 && x) atomicAdd(P.evals + 3, x);
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.evals, (unsigned long long)m * nr);
    }
    if (FUSE_D)
    {
        // every thread publishes its outputs, then one arrival per CTA; the last CTA of the pair continues with kernel D
        __shared__ uint32_t s_last;
        __threadfence();
        __syncthreads();
        if (tid == 0)
        {
            const uint32_t prev = atomicAdd(P.wconst + 2, 1u);
            s_last = (prev + 1u == gridDim.x) ? 1u : 0u;
            if (s_last) P.wconst[2] = 0u;                    // re-armed for the next iteration (nobody else touches it until then)
        }
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        reduce_solve_body<1, THREADS>(P, cfg, 0, 0, reinterpret_cast<float *>(smem_sp4), 0u);
    }
}

// =================================================================================================
// D: reductions in the reference's tree shapes + solve + pose accumulation + loop control.
// One cluster of CL CTAs per pair; partial results travel through global scratch between cluster barriers.
// =================================================================================================
template <typename Tv> __device__ __forceinline__ Tv wtree(Tv a, Tv b, Tv c, Tv d);
template <> __device__ __forceinline__ float wtree<float>(float a, float b, float c, float d) { return warp_tree128(a, b, c, d); }
template <> __device__ __forceinline__ double wtree<double>(double a, double b, double c, double d) { return warp_tree128_d(a, b, c, d); }
template <typename Tv> __device__ __forceinline__ Tv addrn(Tv a, Tv b);
template <> __device__ __forceinline__ float addrn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double addrn<double>(double a, double b) { return __dadd_rn(a, b); }

// rows x cnt values (row stride `stride`) -> one value per row, by levels of 128-slot trees.
// QUAD: slot = ((v0+v1)+v2)+v3 of 4 consecutive values (Reduce<SUM,float>), else slot = value.
// s0/s1: ping-pong scratch (shared memory), row stride sstride >= ceil(cnt/per).  Whole CTA participates.
template <typename Tv, bool QUAD>
__device__ void cta_reduce_rows(const Tv *src, uint32_t rows, uint32_t stride, uint32_t cnt, Tv *s0, Tv *s1, uint32_t sstride, Tv *result)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t per = QUAD ? 512u : 128u;
    const Tv *in = src;
    uint32_t in_stride = stride;
    Tv *out = s0;
    while (true)
    {
        const uint32_t nb = (cnt + per - 1u) / per;
        for (uint32_t item = warp; item < rows * nb; item += nwarps)
        {
            const uint32_t row = item / nb, b = item % nb;
            const Tv *rp = in + (size_t)row * in_stride;
            Tv e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t slot = lane + 32u * j;
                if (QUAD)
                {
                    const uint32_t idx = b * 512u + slot * 4u;
                    const Tv v0 = idx < cnt ? rp[idx] : Tv(0), v1 = idx + 1 < cnt ? rp[idx + 1] : Tv(0);
                    const Tv v2 = idx + 2 < cnt ? rp[idx + 2] : Tv(0), v3 = idx + 3 < cnt ? rp[idx + 3] : Tv(0);
                    e[j] = addrn(addrn(addrn(v0, v1), v2), v3);
                }
                else
                {
                    const uint32_t idx = b * 128u + slot;
                    e[j] = idx < cnt ? rp[idx] : Tv(0);
                }
            }
            const Tv s = wtree<Tv>(e[0], e[1], e[2], e[3]);
            if (lane == 0) out[(size_t)row * sstride + b] = s;
        }
        __syncthreads();
        if (nb == 1)
        {
            if (threadIdx.x < rows) result[threadIdx.x] = out[(size_t)threadIdx.x * sstride];
            __syncthreads();
            return;
        }
        cnt = nb; in = out; in_stride = sstride;
        out = (out == s0) ? s1 : s0;
    }
}

// (float)((double)w / sumw), bit-exact, without a double division per element.
// q' = w * RN(1/sumw) differs from the correctly rounded quotient q by less than 3 ulp (double); (float)q' == (float)q
// unless a float rounding boundary (a midpoint between two floats: low 29 mantissa bits == 0x10000000) lies within that
// distance of q' -- then, and for results outside the normal float range, the exact division is done.
__device__ __forceinline__ float ratio_f32(float w, double sumw, double inv_sumw)
{
    const double qa = __dmul_rn((double)w, inv_sumw);
    const unsigned long long b = (unsigned long long)__double_as_longlong(qa);
    const uint32_t low = (uint32_t)b & 0x1FFFFFFFu;                     // bits below the float mantissa
    const uint32_t ex = (uint32_t)(b >> 52) & 0x7FFu;
    const bool near_boundary = (low - 0x0FFFFFF8u) <= 16u;              // |low - 2^28| <= 8
    const bool normal = ex > 1023u - 126u + 1u && ex < 1023u + 127u;    // float result normal, away from under/overflow
    if (near_boundary || !normal) return (float)__ddiv_rn((double)w, sumw);
    return (float)qa;
}

template <int CL>
__device__ __forceinline__ void cluster_barrier()
{
    if (CL == 1) __syncthreads();
    else cg::this_cluster().sync();
}

#define DRING_DEPTH 3u
#define DRING_STAGE_FLOATS 448u           // 7 arrays x 64 rows
#define DRING_BYTES_T(T) ((size_t)((T) / 32u) * DRING_DEPTH * DRING_STAGE_FLOATS * 4u)        // one ring per warp
#define DRING_BYTES DRING_BYTES_T(512u)
__device__ __forceinline__ void cp_async_wait_2() { asm volatile("cp.async.wait_group 2;" ::: "memory"); }
#define D_SSTRIDE 72u     // scratch row stride in shared memory: supports ceil(cnt/128) <= 72 per level

// The body of kernel D as a device function: called by k_reduce_solve (one CTA / cluster per pair) and, in batch mode, by
// the LAST CTA of a pair to leave k_search_sorted<true> (fused tail: no launch, inputs still hot in L2, and the serial
// solve of one pair overlaps the list scans of the others).  blockDim.x must be T; smem_d: reduce_smem (CL) bytes.
template <int CL, int T>
__device__ __forceinline__ void reduce_solve_body(const PairPtrs &P, const FusedCfg &cfg, cudaGraphConditionalHandle handle, int use_handle,
                                  float *smem_d, const uint32_t rank, const int wide_phase)
{
    // CL = 0 ("wide"): the CTAs of an ordinary grid share the level-1 blocks (CLr = gridDim.x) and the phases are separate
    // launches -- wide_phase 1: block sums of the weights; 2: sum of weights + block means; 3: means + S partials; 4 (one CTA):
    // S, solve, pose.  One large registration keeps every SM busy this way instead of the 16 of a cluster.
    constexpr bool WIDE = (CL == 0);
    const uint32_t CLr = WIDE ? gridDim.x : (uint32_t)CL;
    __shared__ double sh_d[2 * D_SSTRIDE + 8];
    __shared__ float sh_told[8];
    __shared__ float sh_pre[16];                 // {R[9], t[3], s} of the accumulated pose, fetched at the top for solve::accumulate
    __shared__ double sh_sumw;
    __shared__ float sh_mean[8];
    __shared__ float sh_S[12];
    // NOTE: the early exit is uniform over the whole cluster (same flag), so no barrier is left half-populated
    const uint32_t done = __ldcg(&P.state->done);
    const uint32_t m = cfg.m;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr uint32_t NW = T / 32;
    const uint32_t nb128 = (m + 127u) / 128u;
    const uint32_t nb128p = (nb128 + 3u) & ~3u;
    const uint32_t G = (m + 3u) / 4u;
    const uint32_t nb512 = (G + 511u) / 512u;
    float *bs = P.red;                       // [nb128p]
    float *bm = bs + (nb128p + 8u);          // [6][nb128]
    float *sp = bm + (size_t)6 * nb128;      // [11][nb512]
    // shared scratch for the level reductions (float view / double view)
    float *sf0 = smem_d;                     // [11 * D_SSTRIDE]
    float *sf1 = sf0 + 11u * D_SSTRIDE;      // [11 * D_SSTRIDE]
    float *slots = sf1 + 11u * D_SSTRIDE;    // [2 halves][11][128]
    // generic path (batch engine: one CTA per pair; one large registration: 16-CTA cluster): per-warp cp.async rings behind the slots,
    // see phases 2 and 3 (host: cfg.dring)
    const bool use_ring = (T % 512 == 0) && cfg.dring != 0;
    float *wr = slots + 2u * 11u * 128u + warp * (DRING_DEPTH * DRING_STAGE_FLOATS);

    unsigned long long *prof = (rank == 0 && tid == 0) ? P.prof : nullptr;
    if (prof && (!WIDE || wide_phase == 1)) { prof[0] = clock64(); prof[16 + 8 * 3] = gtime_ns(); }
    if (cfg.settle && rank == 0 && tid < 8) sh_told[tid] = __ldcg(P.T + tid);      // consumed by lane 0 of warp 0 after many barriers
    if (rank == 0 && tid < 13) sh_pre[tid] = (tid < 9) ? __ldcg(&P.state->R[tid]) : (tid < 12) ? __ldcg(&P.state->t[tid - 9]) : __ldcg(&P.state->s);
    // Latency-mode fast path (one 8-CTA cluster, m = 8 level-1 blocks of 512 work-items = 16384 points): every CTA loads the
    // 2048 sorted points of ITS level-1 block (4 strided segments of 512) once into shared memory, and the partial
    // results travel through distributed shared memory between cluster barriers -- no global round trip between the
    // phases.  Same slots, same trees, same order as the generic path below => bit-identical.
    bool fast_done = false;
    const bool fast = CL == 8 && cfg.fastD && nb512 == 8u && (m % 2048u) == 0u;
    if (!fast && done) return;
    if (fast)
    {
        cg::cluster_group cluster = cg::this_cluster();
        float *xw = slots + 2u * 11u * 128u;     // [2048] weights of the CTA's points, local index li = j*512 + t
        float *xf = xw + 2048;                   // [3][2048]
        float *xm = xf + 3 * 2048;               // [3][2048]
        float *bs_all = xm + 3 * 2048;           // [128]    level-1 block sums of the weights, every CTA holds all of them
        float *bm_all = bs_all + 128;            // [6][128] level-1 block means
        float *sp_all = bm_all + 6 * 128;        // [11][8]  level-1 S partials (used by rank 0)
        for (uint32_t li = tid; li < 2048u; li += T)
        {
            const uint32_t gp = rank * 512u + (li & 511u) + (li >> 9) * G;
            xw[li] = cfg.weighted ? __ldcg(P.W + gp) : 0.f;
            xf[li] = __ldcg(P.fxyz + gp); xf[2048 + li] = __ldcg(P.fxyz + (size_t)m + gp); xf[4096 + li] = __ldcg(P.fxyz + (size_t)2 * m + gp);
            xm[li] = __ldcg(P.mxyz + gp); xm[2048 + li] = __ldcg(P.mxyz + (size_t)m + gp); xm[4096 + li] = __ldcg(P.mxyz + (size_t)2 * m + gp);
        }
        if (done) return;                        // fetched at the top; uniform over the cluster, nothing written yet
        cluster.sync();                          // data in place; every CTA of the cluster is running (DSMEM may be touched)
        // global 128-point block of local block lb = warp (16 local blocks): 4 per segment
        const uint32_t lb = warp, gb = rank * 4u + (lb & 3u) + (lb >> 2) * 32u;
        double sumw = 1.0;
        if (cfg.weighted)
        {
            if (warp < 16u)
            {
                const float *bp = xw + lb * 128u;
                float sv = warp_tree128(bp[lane], bp[lane + 32], bp[lane + 64], bp[lane + 96]);
                sv = __shfl_sync(FULL_MASK, sv, 0);
                if (lane < (uint32_t)CL) cluster.map_shared_rank(bs_all, lane)[gb] = sv;
            }
            cluster.sync();
            if (warp == 0)
            {
                const float4 q4 = reinterpret_cast<const float4 *>(bs_all)[lane];          // nq = 32 quads of block sums
                const double v = __dadd_rn(__dadd_rn(__dadd_rn((double)q4.x, (double)q4.y), (double)q4.z), (double)q4.w);
                const double sd = warp_tree128_d(v, 0.0, 0.0, 0.0);
                if (lane == 0) sh_sumw = sd;
            }
            __syncthreads();
            sumw = sh_sumw;
            if (rank == 0 && tid == 0) *P.sum_w = sumw;
        }
        if (prof) prof[1] = clock64();
        {
            const float fn = (float)m;
            const double inv_sumw = __ddiv_rn(1.0, sumw);
            if (warp < 16u)
            {
                float e[6][4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t li = lb * 128u + lane + 32u * j;
                    const float fx = xf[li], fy = xf[2048 + li], fz = xf[4096 + li];
                    const float mx = xm[li], my = xm[2048 + li], mz = xm[4096 + li];
                    if (cfg.weighted)
                    {
                        const float wn = ratio_f32(xw[li], sumw, inv_sumw);
                        e[0][j] = __fmul_rn(wn, fx); e[1][j] = __fmul_rn(wn, fy); e[2][j] = __fmul_rn(wn, fz);
                        e[3][j] = __fmul_rn(wn, mx); e[4][j] = __fmul_rn(wn, my); e[5][j] = __fmul_rn(wn, mz);
                    }
                    else
                    {
                        e[0][j] = __fdiv_rn(fx, fn); e[1][j] = __fdiv_rn(fy, fn); e[2][j] = __fdiv_rn(fz, fn);
                        e[3][j] = __fdiv_rn(mx, fn); e[4][j] = __fdiv_rn(my, fn); e[5][j] = __fdiv_rn(mz, fn);
                    }
                }
                float keep = 0.f;                       // lane ch (< 6) ends up with the block mean of channel ch
#pragma unroll
                for (int ch = 0; ch < 6; ++ch)
                {
                    float sv = warp_tree128(e[ch][0], e[ch][1], e[ch][2], e[ch][3]);
                    sv = __shfl_sync(FULL_MASK, sv, 0);
                    if (lane == (uint32_t)ch) keep = sv;
                }
                // 6 channels x 8 peers = 48 remote stores: index = peer * 6 + channel, two rounds of the warp
#pragma unroll
                for (int rd = 0; rd < 2; ++rd)
                {
                    const uint32_t idx = lane + 32u * (uint32_t)rd, peer = idx / 6u, ch = idx % 6u;
                    const float v = __shfl_sync(FULL_MASK, keep, ch);
                    if (idx < 6u * (uint32_t)CL) cluster.map_shared_rank(bm_all, peer)[ch * 128u + gb] = v;
                }
            }
            cluster.sync();
            if (warp < 6u)
            {
                const float *rp = bm_all + warp * 128u;
                const float sv = warp_tree128(rp[lane], rp[lane + 32], rp[lane + 64], rp[lane + 96]);
                if (lane == 0) sh_mean[(warp / 3u) * 4u + warp % 3u] = sv;
            }
            if (tid == 0) { sh_mean[3] = 0.f; sh_mean[7] = 0.f; }
            __syncthreads();
            if (rank == 0 && tid < 8) P.mean[tid] = sh_mean[tid];
        }
        if (prof) prof[2] = clock64();
        {
            const float c = cfg.c;
            const float mfx = sh_mean[0], mfy = sh_mean[1], mfz = sh_mean[2];
            const float mmx = sh_mean[4], mmy = sh_mean[5], mmz = sh_mean[6];
            if (tid < 512u)
            {
                const uint32_t slot_l = tid >> 2, e4 = tid & 3u;
                float A[11];
#pragma unroll
                for (int k = 0; k < 11; ++k) A[k] = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t li = tid + 512u * (uint32_t)j;
                    const float dmx = __fsub_rn(xm[li], mmx), dmy = __fsub_rn(xm[2048 + li], mmy), dmz = __fsub_rn(xm[4096 + li], mmz);
                    const float dfx = __fsub_rn(xf[li], mfx), dfy = __fsub_rn(xf[2048 + li], mfy), dfz = __fsub_rn(xf[4096 + li], mfz);
                    const float mp[3] = { __fmul_rn(c, dmx), __fmul_rn(c, dmy), __fmul_rn(c, dmz) };
                    const float fp[3] = { __fmul_rn(c, dfx), __fmul_rn(c, dfy), __fmul_rn(c, dfz) };
                    const float ff = __fadd_rn(__fadd_rn(__fmul_rn(fp[0], fp[0]), __fmul_rn(fp[1], fp[1])), __fmul_rn(fp[2], fp[2]));
                    const float mm2 = __fadd_rn(__fadd_rn(__fmul_rn(mp[0], mp[0]), __fmul_rn(mp[1], mp[1])), __fmul_rn(mp[2], mp[2]));
                    if (cfg.weighted)
                    {
                        const float w = xw[li];
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int b = 0; b < 3; ++b)
                                A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(w, __fmul_rn(mp[a], fp[b])));
                        A[9] = __fadd_rn(A[9], __fmul_rn(w, ff));
                        A[10] = __fadd_rn(A[10], __fmul_rn(w, mm2));
                    }
                    else
                    {
#pragma unroll
                        for (int a = 0; a < 3; ++a)
#pragma unroll
                            for (int b = 0; b < 3; ++b)
                                A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(mp[a], fp[b]));
                        A[9] = __fadd_rn(A[9], ff);
                        A[10] = __fadd_rn(A[10], mm2);
                    }
                }
#pragma unroll
                for (int k = 0; k < 11; ++k)
                {
                    const float a1 = __shfl_down_sync(FULL_MASK, A[k], 1);
                    const float a2 = __shfl_down_sync(FULL_MASK, A[k], 2);
                    const float a3 = __shfl_down_sync(FULL_MASK, A[k], 3);
                    if (e4 == 0) slots[k * 128 + slot_l] = __fadd_rn(__fadd_rn(__fadd_rn(A[k], a1), a2), a3);
                }
            }
            __syncthreads();
            if (warp < 11u)
            {
                const float *rowp = slots + warp * 128u;
                const float sv = warp_tree128(rowp[lane], rowp[lane + 32], rowp[lane + 64], rowp[lane + 96]);
                if (lane == 0) cluster.map_shared_rank(sp_all, 0)[warp * 8u + rank] = sv;
            }
            cluster.sync();
            if (rank != 0) return;
            if (prof) prof[3] = clock64();
            // second level of reduce_sum_f over the 8 block results of every row: slots 0 and 1 hold a quad each
            if (warp < 11u)
            {
                const float *rp = sp_all + warp * 8u;
                float e0 = 0.f;
                if (lane < 2u) e0 = __fadd_rn(__fadd_rn(__fadd_rn(rp[lane * 4u], rp[lane * 4u + 1u]), rp[lane * 4u + 2u]), rp[lane * 4u + 3u]);
                const float sv = warp_tree128(e0, 0.f, 0.f, 0.f);
                if (lane == 0) sh_S[warp] = sv;
            }
            __syncthreads();
        }
        fast_done = true;
    }
    if (!fast_done)
    {
    // ---------------- phase 1: sum of weights (ICPWeights) ----------------
    double sumw = 1.0;
    if (cfg.weighted && (!WIDE || wide_phase <= 2))
    {
        if (!WIDE || wide_phase == 1)
        // 8 blocks of a warp at a time: all 32 loads of a lane are issued before the first tree (the plain loop exposed one memory
        // latency per block: 10.8 K cycles for 8 blocks per warp in the batch engine's tail, tools/d_phases.py)
        for (uint32_t base = rank * NW + warp; base < nb128p; base += 8u * CLr * NW)
        {
            float e[8][4];
#pragma unroll
            for (uint32_t u = 0; u < 8u; ++u)
            {
                const uint32_t blk = base + u * CLr * NW;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t idx = blk * 128u + lane + 32u * j;
                    e[u][j] = (blk < nb128p && idx < m) ? __ldcg(P.W + idx) : 0.f;
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < 8u; ++u)
            {
                const uint32_t blk = base + u * CLr * NW;
                const float s = warp_tree128(e[u][0], e[u][1], e[u][2], e[u][3]);
                if (lane == 0 && blk < nb128p) bs[blk] = s;
            }
        }
        if (WIDE && wide_phase == 1) return;
        if (!WIDE) cluster_barrier<CL>();
        // every CTA finishes the sum redundantly (identical operations => identical value)
        if (nb128 == 1) { if (tid == 0) sh_sumw = (double)__ldcg(bs); __syncthreads(); }
        else
        {
            const uint32_t nq = nb128p / 4u;
            // quads of block sums -> f64 (reduce_sum_fd), then 128-slot levels.  nq <= 128*D_SSTRIDE guaranteed by init.
            // First level straight from global memory into sh_d (ceil(nq/128) values).
            const uint32_t nbq = (nq + 127u) / 128u;
            for (uint32_t b = warp; b < nbq; b += NW)
            {
                double e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t k = b * 128u + lane + 32u * j;
                    double v = 0.0;
                    if (k < nq)
                    {
                        const float4 q4 = __ldcg((const float4 *)bs + k);
                        v = __dadd_rn(__dadd_rn(__dadd_rn((double)q4.x, (double)q4.y), (double)q4.z), (double)q4.w);
                    }
                    e[j] = v;
                }
                const double s = warp_tree128_d(e[0], e[1], e[2], e[3]);
                if (lane == 0) sh_d[b] = s;
            }
            __syncthreads();
            if (nbq == 1) { if (tid == 0) sh_sumw = sh_d[0]; __syncthreads(); }
            else cta_reduce_rows<double, false>(sh_d, 1, 0, nbq, sh_d + D_SSTRIDE, sh_d + D_SSTRIDE + 4, 0, &sh_sumw);
        }
        sumw = sh_sumw;
        if (rank == 0 && tid == 0) *P.sum_w = sumw;
    }

    if (WIDE && wide_phase == 1) return;               // unweighted: nothing to do in the first launch
    if (prof) prof[1] = clock64();
    // ---------------- phase 2: (weighted) means (ICPMean) ----------------
    {
        const float fn = (float)m;
        const double inv_sumw = __ddiv_rn(1.0, sumw);
        if (WIDE && wide_phase != 2) { }
        else if (use_ring)
        {
            // every warp streams its blocks through a private 3-deep ring of half blocks (64 rows x 7 arrays = 1792 bytes),
            // filled by 16-byte cp.async copies: two half blocks are in flight while one is consumed, no registers held
            const uint32_t blk0 = rank * NW + warp, bstride = CLr * NW;            // this warp's blocks: blk0, blk0 + bstride, ...
            const uint32_t nblk_w = blk0 < nb128 ? (nb128 - blk0 + bstride - 1u) / bstride : 0u, nhs = 2u * nblk_w;
            auto issue = [&](uint32_t hs) {
                if (hs < nhs)
                {
                    const uint32_t r0 = (blk0 + bstride * (hs >> 1)) * 128u + (hs & 1u) * 64u;
                    float *stage = wr + (hs % DRING_DEPTH) * DRING_STAGE_FLOATS;
#pragma unroll
                    for (uint32_t t4 = 0; t4 < 4u; ++t4)
                    {
                        const uint32_t gi = lane + 32u * t4;
                        if (gi < 112u)
                        {
                            const uint32_t a = gi >> 4, gr = gi & 15u;
                            const float *src = (a == 0u) ? P.W : (a < 4u ? P.fxyz + (size_t)(a - 1u) * m : P.mxyz + (size_t)(a - 4u) * m);
                            cp_async16(stage + a * 64u + gr * 4u, src + r0 + gr * 4u);
                        }
                    }
                }
                cp_async_commit();
            };
            issue(0); issue(1); issue(2);
            for (uint32_t b = 0; b < nblk_w; ++b)
            {
                float e[6][4];
#pragma unroll
                for (uint32_t h = 0; h < 2u; ++h)
                {
                    const uint32_t hs = 2u * b + h;
                    cp_async_wait_2();
                    __syncwarp();
                    const float *S = wr + (hs % DRING_DEPTH) * DRING_STAGE_FLOATS;
#pragma unroll
                    for (uint32_t jj = 0; jj < 2u; ++jj)
                    {
                        const uint32_t li = lane + 32u * jj;
                        const float fx = S[64u + li], fy = S[128u + li], fz = S[192u + li];
                        const float mx = S[256u + li], my = S[320u + li], mz = S[384u + li];
                        float v[6];
                        if (cfg.weighted)
                        {
                            const float wn = ratio_f32(S[li], sumw, inv_sumw);
                            v[0] = __fmul_rn(wn, fx); v[1] = __fmul_rn(wn, fy); v[2] = __fmul_rn(wn, fz);
                            v[3] = __fmul_rn(wn, mx); v[4] = __fmul_rn(wn, my); v[5] = __fmul_rn(wn, mz);
                        }
                        else
                        {
                            v[0] = __fdiv_rn(fx, fn); v[1] = __fdiv_rn(fy, fn); v[2] = __fdiv_rn(fz, fn);
                            v[3] = __fdiv_rn(mx, fn); v[4] = __fdiv_rn(my, fn); v[5] = __fdiv_rn(mz, fn);
                        }
#pragma unroll
                        for (int ch = 0; ch < 6; ++ch) e[ch][2u * h + jj] = v[ch];
                    }
                    __syncwarp();
                    issue(hs + DRING_DEPTH);
                }
                const uint32_t blk = blk0 + bstride * b;
#pragma unroll
                for (int ch = 0; ch < 6; ++ch)
                {
                    const float sv = warp_tree128(e[ch][0], e[ch][1], e[ch][2], e[ch][3]);
                    if (lane == 0) bm[(size_t)ch * nb128 + blk] = sv;
                }
            }
            cp_async_wait_all();
        }
        else
        for (uint32_t blk = rank * NW + warp; blk < nb128; blk += CLr * NW)
        {
            float e[6][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t idx = blk * 128u + lane + 32u * j;
                float v[6] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
                if (idx < m)
                {
                    float fx = __ldcg(P.fxyz + idx), fy = __ldcg(P.fxyz + (size_t)m + idx), fz = __ldcg(P.fxyz + (size_t)2 * m + idx);
                    float mx = __ldcg(P.mxyz + idx), my = __ldcg(P.mxyz + (size_t)m + idx), mz = __ldcg(P.mxyz + (size_t)2 * m + idx);
                    if (cfg.weighted)
                    {
                        const float wn = ratio_f32(__ldcg(P.W + idx), sumw, inv_sumw);
                        v[0] = __fmul_rn(wn, fx); v[1] = __fmul_rn(wn, fy); v[2] = __fmul_rn(wn, fz);
                        v[3] = __fmul_rn(wn, mx); v[4] = __fmul_rn(wn, my); v[5] = __fmul_rn(wn, mz);
                    }
                    else
                    {
                        v[0] = __fdiv_rn(fx, fn); v[1] = __fdiv_rn(fy, fn); v[2] = __fdiv_rn(fz, fn);
                        v[3] = __fdiv_rn(mx, fn); v[4] = __fdiv_rn(my, fn); v[5] = __fdiv_rn(mz, fn);
                    }
                }
#pragma unroll
                for (int ch = 0; ch < 6; ++ch) e[ch][j] = v[ch];
            }
#pragma unroll
            for (int ch = 0; ch < 6; ++ch)
            {
                const float s = warp_tree128(e[ch][0], e[ch][1], e[ch][2], e[ch][3]);
                if (lane == 0) bm[(size_t)ch * nb128 + blk] = s;
            }
        }
        if (WIDE && wide_phase == 2) return;
        if (!WIDE) cluster_barrier<CL>();
        if (nb128 == 1)
        {
            if (tid < 6) sh_mean[(tid / 3u) * 4u + tid % 3u] = __ldcg(bm + tid);
            __syncthreads();
        }
        else
        {
            // first level from global (volatile-ish loads), next levels in shared memory
            const uint32_t nb2 = (nb128 + 127u) / 128u;
            for (uint32_t item = warp; item < 6u * nb2; item += NW)
            {
                const uint32_t ch = item / nb2, b = item % nb2;
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                {
                    const uint32_t idx = b * 128u + lane + 32u * j;
                    e[j] = idx < nb128 ? __ldcg(bm + (size_t)ch * nb128 + idx) : 0.f;
                }
                const float s = warp_tree128(e[0], e[1], e[2], e[3]);
                if (lane == 0) sf0[ch * D_SSTRIDE + b] = s;
            }
            __syncthreads();
            if (nb2 == 1) { if (tid < 6) sh_mean[(tid / 3u) * 4u + tid % 3u] = sf0[tid * D_SSTRIDE]; __syncthreads(); }
            else
            {
                __shared__ float tmp6[8];
                cta_reduce_rows<float, false>(sf0, 6, D_SSTRIDE, nb2, sf1, sf0 + 6u * D_SSTRIDE, D_SSTRIDE / 2u, tmp6);
                if (tid < 6) sh_mean[(tid / 3u) * 4u + tid % 3u] = tmp6[tid];
                __syncthreads();
            }
        }
        if (tid == 0) { sh_mean[3] = 0.f; sh_mean[7] = 0.f; }
        __syncthreads();
        if (rank == 0 && tid < 8) P.mean[tid] = sh_mean[tid];
    }

    if (prof) prof[2] = clock64();
    // ---------------- phase 3: deviations + S_ij partial sums (ICPDevs + ICPS) ----------------
    {
        const float c = cfg.c;
        const float mfx = sh_mean[0], mfy = sh_mean[1], mfz = sh_mean[2];
        const float mmx = sh_mean[4], mmy = sh_mean[5], mmz = sh_mean[6];
        // one level-1 block = 512 work-items of the reference kernel = 128 slots of 4 consecutive work-items.
        // 4 adjacent lanes own one slot (one work-item each).  A CTA of T threads works on NH = max(T/512, 1) blocks at a
        // time, each by TH = min(T, 512) threads in SP = 512/TH passes of TH/4 slots.
        constexpr uint32_t TH = (T >= 512) ? 512u : (uint32_t)T, NH = (T >= 512) ? (uint32_t)T / 512u : 1u, SP = 512u / TH;
        constexpr uint32_t WH = TH / 32u;                       // warps per block group
        const uint32_t half = tid / TH, th = tid % TH, e = th & 3u;
        const uint32_t nrounds = (nb512 + CLr * NH - 1u) / (CLr * NH);
        if (WIDE && wide_phase != 3) { }
        else if (use_ring)
        {
            // same ring: a warp's 32 work-items of block B need 4 strided segments of 32 rows; half stage = segments {2 jh, 2 jh + 1}.
            // Round rd: block group `half` of the CTA works on block B = (rd * NH + half) * CLr + rank (as in the plain loop).
            const uint32_t nhs = 2u * nrounds, wl = warp % WH;
            auto issue = [&](uint32_t hs) {
                const uint32_t B = ((hs >> 1) * NH + half) * CLr + rank;
                if (hs < nhs && B < nb512)
                {
                    const uint32_t g0 = B * 512u + 32u * wl, j0 = 2u * (hs & 1u);
                    float *stage = wr + (hs % DRING_DEPTH) * DRING_STAGE_FLOATS;
#pragma unroll
                    for (uint32_t t4 = 0; t4 < 4u; ++t4)
                    {
                        const uint32_t gi = lane + 32u * t4;
                        if (gi < 112u)
                        {
                            const uint32_t a = gi >> 4, jj = (gi >> 3) & 1u, gr = gi & 7u;
                            const float *src = (a == 0u) ? P.W : (a < 4u ? P.fxyz + (size_t)(a - 1u) * m : P.mxyz + (size_t)(a - 4u) * m);
                            cp_async16(stage + a * 64u + jj * 32u + gr * 4u, src + g0 + (size_t)(j0 + jj) * G + gr * 4u);
                        }
                    }
                }
                cp_async_commit();
            };
            issue(0); issue(1); issue(2);
            float *gs = slots + (size_t)half * 11u * 128u;
            const uint32_t slot_l = th >> 2;
            for (uint32_t rd = 0; rd < nrounds; ++rd)
            {
                const uint32_t B = (rd * NH + half) * CLr + rank;
                float A[11];
#pragma unroll
                for (int k = 0; k < 11; ++k) A[k] = 0.f;
#pragma unroll
                for (uint32_t jh = 0; jh < 2u; ++jh)
                {
                    const uint32_t hs = 2u * rd + jh;
                    cp_async_wait_2();
                    __syncwarp();
                    const float *S = wr + (hs % DRING_DEPTH) * DRING_STAGE_FLOATS;
                    if (B < nb512)
                    {
#pragma unroll
                    for (uint32_t jj = 0; jj < 2u; ++jj)
                    {
                        const uint32_t li = jj * 32u + lane;
                        const float dmx = __fsub_rn(S[256u + li], mmx), dmy = __fsub_rn(S[320u + li], mmy), dmz = __fsub_rn(S[384u + li], mmz);
                        const float dfx = __fsub_rn(S[64u + li], mfx), dfy = __fsub_rn(S[128u + li], mfy), dfz = __fsub_rn(S[192u + li], mfz);
                        const float mp[3] = { __fmul_rn(c, dmx), __fmul_rn(c, dmy), __fmul_rn(c, dmz) };
                        const float fp[3] = { __fmul_rn(c, dfx), __fmul_rn(c, dfy), __fmul_rn(c, dfz) };
                        const float ff = __fadd_rn(__fadd_rn(__fmul_rn(fp[0], fp[0]), __fmul_rn(fp[1], fp[1])), __fmul_rn(fp[2], fp[2]));
                        const float mm2 = __fadd_rn(__fadd_rn(__fmul_rn(mp[0], mp[0]), __fmul_rn(mp[1], mp[1])), __fmul_rn(mp[2], mp[2]));
                        if (cfg.weighted)
                        {
                            const float w = S[li];
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int b = 0; b < 3; ++b)
                                    A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(w, __fmul_rn(mp[a], fp[b])));
                            A[9] = __fadd_rn(A[9], __fmul_rn(w, ff));
                            A[10] = __fadd_rn(A[10], __fmul_rn(w, mm2));
                        }
                        else
                        {
#pragma unroll
                            for (int a = 0; a < 3; ++a)
#pragma unroll
                                for (int b = 0; b < 3; ++b)
                                    A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(mp[a], fp[b]));
                            A[9] = __fadd_rn(A[9], ff);
                            A[10] = __fadd_rn(A[10], mm2);
                        }
                    }
                    }
                    __syncwarp();
                    issue(hs + DRING_DEPTH);
                }
#pragma unroll
                for (int k = 0; k < 11; ++k)
                {
                    const float a1 = __shfl_down_sync(FULL_MASK, A[k], 1);
                    const float a2 = __shfl_down_sync(FULL_MASK, A[k], 2);
                    const float a3 = __shfl_down_sync(FULL_MASK, A[k], 3);
                    if (e == 0) gs[k * 128 + slot_l] = __fadd_rn(__fadd_rn(__fadd_rn(A[k], a1), a2), a3);
                }
                __syncthreads();
                if (B < nb512)
                {
                    for (uint32_t k = wl; k < 11u; k += WH)
                    {
                        const float *rowp = gs + k * 128u;
                        const float sv = warp_tree128(rowp[lane], rowp[lane + 32], rowp[lane + 64], rowp[lane + 96]);
                        if (lane == 0) sp[(size_t)k * nb512 + B] = sv;
                    }
                }
                __syncthreads();
            }
            cp_async_wait_all();
        }
        else
        for (uint32_t rd = 0; rd < nrounds; ++rd)
        {
            const uint32_t B = (rd * NH + half) * CLr + rank;     // blocks interleaved over the cluster
            float *gs = slots + (size_t)half * 11u * 128u;
#pragma unroll 1
            for (uint32_t pass = 0; pass < SP; ++pass)
            {
                const uint32_t slot_l = pass * (TH / 4u) + (th >> 2);
                float A[11];
#pragma unroll
                for (int k = 0; k < 11; ++k) A[k] = 0.f;
                const uint32_t g = (B * 128u + slot_l) * 4u + e;      // work-item of the reference kernel
                if (B < nb512 && g < G)
                {
                    // the (up to) 4 strided pairs of the work-item: issue all loads first, then accumulate in order
                    float w4[4], f4[4][3], m4[4][3];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const uint32_t pi = g + (uint32_t)j * G;
                        const bool ok = pi < m;
                        const uint32_t pj = ok ? pi : g;
                        w4[j] = __ldcg(P.W + pj);
                        f4[j][0] = __ldcg(P.fxyz + pj); f4[j][1] = __ldcg(P.fxyz + (size_t)m + pj); f4[j][2] = __ldcg(P.fxyz + (size_t)2 * m + pj);
                        m4[j][0] = __ldcg(P.mxyz + pj); m4[j][1] = __ldcg(P.mxyz + (size_t)m + pj); m4[j][2] = __ldcg(P.mxyz + (size_t)2 * m + pj);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                    {
                        const uint32_t pi = g + (uint32_t)j * G;
                        if (pi < m)
                        {
                            const float dmx = __fsub_rn(m4[j][0], mmx), dmy = __fsub_rn(m4[j][1], mmy), dmz = __fsub_rn(m4[j][2], mmz);
                            const float dfx = __fsub_rn(f4[j][0], mfx), dfy = __fsub_rn(f4[j][1], mfy), dfz = __fsub_rn(f4[j][2], mfz);
                            const float mp[3] = { __fmul_rn(c, dmx), __fmul_rn(c, dmy), __fmul_rn(c, dmz) };
                            const float fp[3] = { __fmul_rn(c, dfx), __fmul_rn(c, dfy), __fmul_rn(c, dfz) };
                            const float ff = __fadd_rn(__fadd_rn(__fmul_rn(fp[0], fp[0]), __fmul_rn(fp[1], fp[1])), __fmul_rn(fp[2], fp[2]));
                            const float mm2 = __fadd_rn(__fadd_rn(__fmul_rn(mp[0], mp[0]), __fmul_rn(mp[1], mp[1])), __fmul_rn(mp[2], mp[2]));
                            if (cfg.weighted)
                            {
                                const float w = w4[j];
#pragma unroll
                                for (int a = 0; a < 3; ++a)
#pragma unroll
                                    for (int b = 0; b < 3; ++b)
                                        A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(w, __fmul_rn(mp[a], fp[b])));
                                A[9] = __fadd_rn(A[9], __fmul_rn(w, ff));
                                A[10] = __fadd_rn(A[10], __fmul_rn(w, mm2));
                            }
                            else
                            {
#pragma unroll
                                for (int a = 0; a < 3; ++a)
#pragma unroll
                                    for (int b = 0; b < 3; ++b)
                                        A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(mp[a], fp[b]));
                                A[9] = __fadd_rn(A[9], ff);
                                A[10] = __fadd_rn(A[10], mm2);
                            }
                        }
                    }
                }
                // slot = ((A_e0 + A_e1) + A_e2) + A_e3 over the 4 lanes of the slot (reduce_sum_f: dot (float4, 1.f))
#pragma unroll
                for (int k = 0; k < 11; ++k)
                {
                    const float a1 = __shfl_down_sync(FULL_MASK, A[k], 1);
                    const float a2 = __shfl_down_sync(FULL_MASK, A[k], 2);
                    const float a3 = __shfl_down_sync(FULL_MASK, A[k], 3);
                    if (e == 0) gs[k * 128 + slot_l] = __fadd_rn(__fadd_rn(__fadd_rn(A[k], a1), a2), a3);
                }
            }
            __syncthreads();
            if (B < nb512)
            {
                // the WH warps of the block group share the 11 rows
                for (uint32_t k = (warp % WH); k < 11u; k += WH)
                {
                    const float *rowp = gs + k * 128u;
                    const float sv = warp_tree128(rowp[lane], rowp[lane + 32], rowp[lane + 64], rowp[lane + 96]);
                    if (lane == 0) sp[(size_t)k * nb512 + B] = sv;
                }
            }
            __syncthreads();
        }
    }
    if (WIDE && wide_phase == 3) return;
    if (!WIDE) cluster_barrier<CL>();
    if (rank != 0) return;
    if (prof) prof[3] = clock64();

    // ---------------- phase 4: second reduction level, solve, pose update ----------------
    if (nb512 == 1) { if (tid < 11) sh_S[tid] = __ldcg(sp + tid); __syncthreads(); }
    else
    {
        // copy the group sums to shared memory (nb512 <= 2048 supported), then quad levels
        // first level straight from global memory
        const uint32_t nb2 = (nb512 + 511u) / 512u;
        for (uint32_t item = warp; item < 11u * nb2; item += NW)
        {
            const uint32_t k = item / nb2, b = item % nb2;
            const float *rp = sp + (size_t)k * nb512;
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t idx = b * 512u + (lane + 32u * j) * 4u;
                const float v0 = idx < nb512 ? __ldcg(rp + idx) : 0.f, v1 = idx + 1 < nb512 ? __ldcg(rp + idx + 1) : 0.f;
                const float v2 = idx + 2 < nb512 ? __ldcg(rp + idx + 2) : 0.f, v3 = idx + 3 < nb512 ? __ldcg(rp + idx + 3) : 0.f;
                e[j] = __fadd_rn(__fadd_rn(__fadd_rn(v0, v1), v2), v3);
            }
            const float s = warp_tree128(e[0], e[1], e[2], e[3]);
            if (lane == 0) sf0[k * D_SSTRIDE + b] = s;
        }
        __syncthreads();
        if (nb2 == 1) { if (tid < 11) sh_S[tid] = sf0[tid * D_SSTRIDE]; __syncthreads(); }
        else cta_reduce_rows<float, true>(sf0, 11, D_SSTRIDE, nb2, sf1, sf1 + 11u * 4u, 4u, sh_S);   // m > 2^20: rejected by init
    }
    }
    if (warp == 0)
    {
        __shared__ __align__(16) float pm_ring[32][4];
        float s11[11], mu[8], tk[8], rk[9], t8[8];
        for (int i = 0; i < 11; ++i) s11[i] = sh_S[i];
        for (int i = 0; i < 8; ++i) mu[i] = sh_mean[i];
        if (lane == 0) for (int i = 0; i < 11; ++i) P.S[i] = s11[i];
        if (prof) prof[4] = clock64();
        if (cfg.power_method)
        {
            // the whole warp runs the power method (see power_method_warp2); lane 0 publishes
            const int pm_iters = solve::power_method_warp2(s11, mu, tk, pm_ring);
            if (prof) { prof[7] = (unsigned long long)pm_iters; prof[6] = clock64(); }
            if (lane == 0) solve::accumulate(P.state, tk, nullptr, t8, sh_pre);
        }
        else if (lane == 0)
        {
            solve::svd_solve(s11, mu, tk, rk);
            for (int i = 0; i < 9; ++i) P.Rk[i] = rk[i];
            solve::accumulate(P.state, tk, rk, t8, sh_pre);
        }
        if (lane == 0)
        {
            if (cfg.settle)
            {
                // the pose the search of THIS iteration used stays available to the next one (settle tests of kernels A and C');
                // it was fetched into shared memory at the top of the kernel, off the critical path
                float *Tprev = reinterpret_cast<float *>(P.wconst + 4);
                for (int i = 0; i < 8; ++i) Tprev[i] = sh_told[i];
                P.wconst[13] = 1u;               // the bounds recorded by this iteration's kernels A / C' belong to the current moving set
            }
            for (int i = 0; i < 8; ++i) { P.Tk[i] = tk[i]; P.T[i] = t8[i]; }
            LoopParams *lp = P.loop;
            const int left = lp->iters_left - 1;
            lp->iters_left = left;
            unsigned cont;
            if (lp->check)
            {
                solve::check_convergence(P.state, lp->max_iterations, lp->angle_thr, lp->trans_thr);
                cont = (P.state->done == 0u && left > 0) ? 1u : 0u;
            }
            else
            {
                P.state->k = P.state->k + 1;
                cont = left > 0 ? 1u : 0u;
            }
            if (use_handle) cudaGraphSetConditional(handle, cont);
            if (prof) { prof[5] = clock64(); prof[16 + 8 * 3 + 6] = gtime_ns(); }
        }
    }
}

// =================================================================================================
// host side
// =================================================================================================
static size_t assign_smem_bytes(uint32_t nr, uint32_t QB, int par_rank)
{
    return (size_t)nr * 32 + (size_t)QB * 4 + (size_t)nr * 4 + (par_rank ? (size_t)div_up(QB, 32) * nr * 2 + 4 : 0);
}

static size_t grouped_smem_bytes(const FusedCfg &cfg);

void fused_choose_cfg(FusedCfg *cfg, uint32_t m, uint32_t nr, int sm_count, uint32_t n_pairs)
{
    cfg->m = m; cfg->nr = nr;
    cfg->lm_w = 0u; cfg->lm_h = 0u; cfg->ctile_w = 0u;       // the single engine fills these in (fused_cfg_of)
    const uint64_t total = (uint64_t)m * n_pairs;
    uint32_t QB;
    if (total <= (uint64_t)sm_count * 1024u)
    {
        // latency mode: one chunk per SM, all of its points in flight at once
        QB = div_up(m, (uint32_t)sm_count);
        QB = (QB + 3u) & ~3u;
        if (QB > 256u) QB = 256u;         // one pass of the 256-thread CTA; beyond one chunk per SM more, shorter CTAs win (65536 points:
                                          // 444-point chunks 93.3 / 106.3 us per iteration at 512 / 1024 representatives, 256-point chunks 86.3 / 95.6)
        if (QB < 32u) QB = 32u;
    }
    else QB = 1024u;         // batch mode (tools/tune2.py sweep with the pruned kernel A; 512 for the exhaustive one)
    // ONE large registration (e.g. 307200 landmarks): 1024-point chunks would give 300 CTAs for 592 resident slots -- half a
    // wave (ncu launch list, profiles/r02_scaled_307200_1024.md).  512-point chunks fill the machine.
    const bool big_single = n_pairs == 1u && total > (uint64_t)sm_count * 1024u;
    if (big_single && assign_smem_bytes(nr, 512u, 1) <= 96u * 1024u) QB = 512u;
    const bool batch = total > (uint64_t)sm_count * 1024u;
    uint32_t TPB = batch ? 512u : 1024u;
    int QPT = batch ? 4 : 2;
    if (const char *e = getenv("ICP_B200_QB")) { int v = atoi(e); if (v >= 32 && v <= 1024 && v % 4 == 0) QB = (uint32_t)v; }
    if (const char *e = getenv("ICP_B200_TPB")) { int v = atoi(e); if (v == 256 || v == 512 || v == 1024) TPB = (uint32_t)v; }
    if (const char *e = getenv("ICP_B200_QPT")) { int v = atoi(e); if (v == 2 || v == 4) QPT = v; }
    if (QPT == 4 && TPB > 512u) TPB = 512u;      // the QPT = 4 instantiation is compiled for <= 512 threads (128 registers)
    // lanes per point group: the largest power of two that still covers the chunk in one pass
    // (TPB / S groups x QPT points >= QB)
    int S = 1;
    while (S < 32 && (TPB / (uint32_t)(S * 2)) * (uint32_t)QPT >= QB) S <<= 1;
    while ((uint32_t)S > nr) S >>= 1;
    // experiment knobs (tools/tune.py); results are independent of them
    if (const char *e = getenv("ICP_B200_S")) { int v = atoi(e); if ((v & (v - 1)) == 0 && v >= 1 && v <= 32 && (uint32_t)v <= nr) S = v; }
    cfg->TPB = TPB; cfg->QPT = QPT;
    cfg->QB = QB; cfg->S = S;
    cfg->nbA = div_up(m, QB);
    cfg->par_rank = (assign_smem_bytes(nr, QB, 1) <= 96u * 1024u) ? 1 : 0;
    if (const char *e = getenv("ICP_B200_PAR_RANK")) { if (atoi(e) == 0) cfg->par_rank = 0; }
    // kernel A flavour: triangle-inequality pruning needs >= 2 stored neighbours per representative
    cfg->K = ((nr - 1u < FUSED_NBR_K ? nr - 1u : FUSED_NBR_K)) & ~1u;
    cfg->Amode = (cfg->K >= 2u && nr <= 4096u) ? 1 : 0;
    if (const char *e = getenv("ICP_B200_AMODE")) { int v = atoi(e); if (v == 0) cfg->Amode = 0; }
    if (cfg->Amode == 1)
    {
        // one point per lane: a CTA of min(512, QB rounded up to a warp) threads covers the chunk in ceil(QB / TPB) passes
        uint32_t t = (QB + 31u) & ~31u;
        if (t > 256u) t = 256u;                    // batch mode: 256-thread CTAs, 4 resident per SM overlap each other's latency phases
        if (!batch && t < 256u) t = 256u;          // latency mode: more lanes for the full-scan pass of the unsettled points
        cfg->TPB = t;
        if (const char *e = getenv("ICP_B200_TPB")) { int v = atoi(e); if (v == 128 || v == 256 || v == 512) cfg->TPB = (uint32_t)v; }
    }
    cfg->SF = batch ? 8 : 32;
    if (const char *e = getenv("ICP_B200_SF")) { int v = atoi(e); if (v == 8 || v == 9 || v == 16 || v == 32) cfg->SF = v; }   // 9 = 8 lanes, independent evaluations (scan_reps_ilp)
    cfg->CL = (n_pairs * 8u <= (uint32_t)sm_count) ? 8 : 1;
    // one large registration: kernel D's generic path strides its blocks over the cluster -- 16 CTAs (non-portable cluster size)
    // halve it (94.7 us of a 377 us iteration at 307200 / 1024 with 8)
    if (n_pairs == 1u && m >= 65536u) cfg->CL = 16;
    if (const char *e = getenv("ICP_B200_CL")) { int v = atoi(e); if (n_pairs == 1u && (v == 8 || v == 16)) cfg->CL = v; }
    // ... and beyond ~10^5 points the wide flavour (whole-GPU launches per pass, k_reduce_wide) replaces the cluster altogether
    // (measured: 307200 / 1024 320.7 -> 302.8 us per iteration, 307200 / 512 399.8 -> 384.0; at 65536 points the three extra launches
    // cost more than the 16-SM cluster loses: 92.9 -> 97.5 us, hence the threshold)
    cfg->wideD = (n_pairs == 1u && m >= 131072u) ? 1 : 0;
    if (const char *e = getenv("ICP_B200_WIDED")) cfg->wideD = (atoi(e) != 0 && n_pairs == 1u) ? 1 : 0;
    cfg->fastD = 1;
    if (const char *e = getenv("ICP_B200_FASTD")) { if (atoi(e) == 0) cfg->fastD = 0; }
    cfg->TD = (cfg->CL >= 8) ? 1024 : 512;      // batch: 2 resident CTAs per SM => 256 pairs fit one wave (tools/gpu_quick.sh sweep)
    if (const char *e = getenv("ICP_B200_TD")) { int v = atoi(e); if (cfg->CL == 1 && (v == 256 || v == 512 || v == 1024)) cfg->TD = v; }
    cfg->L = 8;
    // queries per CTA in kernel C: enough CTAs to cover the SMs in latency mode, amortised prologue in batch mode
    cfg->QC = (total <= (uint64_t)sm_count * 1024u) ? 32u : 512u;
    if (total > (uint64_t)sm_count * 1024u) cfg->L = 4;
    if (const char *e = getenv("ICP_B200_L")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) cfg->L = v; }
    if (const char *e = getenv("ICP_B200_QC")) { int v = atoi(e); if (v >= 32 && v % 32 == 0) cfg->QC = (uint32_t)v; }
    // grouped kernel C: batch mode = 1024-query CTAs, full-warp items; latency mode = one A-chunk per CTA, 8-query items
    const bool batch_mode = total > (uint64_t)sm_count * 1024u;
    cfg->Cmode = 1;
    // queries per CTA of the grouped kernel C (independent of kernel A's chunks): 1024 in batch mode, one slice per SM in latency mode
    {
        uint32_t qg = batch_mode ? (big_single ? 512u : 1024u) : ((div_up(m, 2u * (uint32_t)sm_count) + 3u) & ~3u);     // latency mode: 2 CTAs per SM balance the skewed lists
        if (qg < 32u) qg = 32u;
        if (qg > 2048u) qg = 2048u;
        cfg->QG = qg;
    }
    cfg->QI = batch_mode ? 32u : 8u;
    // batch mode: the sorted flavour (B' sorts, C' owns 2048 consecutive sorted positions) when the chunk histograms fit shared memory
    const bool sort_fits = ((size_t)cfg->nbA * nr + 6u * nr) * 4u + 16u <= 200u * 1024u && nr <= 65535u;
    if (batch_mode && sort_fits && (size_t)cfg->nbA * nr * 4u <= 96u * 1024u) { cfg->Cmode = 2; cfg->QG = 2048u; }
    // The span flavour (Cmode 3: sorted query records from B'', the CTA's list span staged in shared memory by cp.async.bulk) is
    // built, parity-tested and MEASURED SLOWER on B200 (profiles/r02_ab_span_vs_sorted.md: B'' 0.123 + C'' 0.348 ms per 256-pair
    // launch against B' 0.023 + C' 0.323): moving 48 bytes per query into sorted order costs more HBM traffic than the gathers it
    // replaces, and the list scans were not tile-staging bound.  It stays selectable (ICP_B200_CMODE=3) for that A/B.
    if (const char *e = getenv("ICP_B200_CMODE")) { int v = atoi(e); if (v == 0 || v == 1 || ((v == 2 || v == 3) && sort_fits)) { cfg->Cmode = v; if (v >= 2) cfg->QG = 2048u; } }
    // sorted flavour: CTAs of B' (each scans all columns, scatters its slice of the queries) and threads per CTA of C'
    // exact temporal pruning of stage 2: batch engine only (its poses change only through kernel D), metric weights in [0, 1]
    cfg->settle = (batch_mode && cfg->Cmode >= 2) ? 1 : 0;
    // ... and in the grouped kernel C (single registrations): the engine distrusts the bounds at the start of every run call
    // (PairPtrs::wconst[13]), since the caller may have replaced the moving set in between
    // Measured (tools/scaled_ab.py, us per iteration with / without): 65536/512 95.9 / 109.5, 65536/1024 108.8 / 128.9,
    // 307200/512 423 / 521, 307200/1024 377 / 444 -- but 16384/256 48.0 / 45.3: in latency mode the extra gathers of the test
    // cost more than the scans it saves.  So: single registrations of >= 32768 points.
    if (cfg->Cmode == 1 && cfg->Amode == 1 && n_pairs == 1u && m >= 32768u) cfg->settle = 1;
    if (const char *e = getenv("ICP_B200_SETTLE")) { cfg->settle = (atoi(e) != 0 && cfg->Cmode >= 1 && cfg->Amode == 1) ? 1 : 0; }
    cfg->aperm = batch_mode ? 1 : 0;     // kernel A hands the chunk's points to the lanes grouped by last iteration's representative
    if (const char *e = getenv("ICP_B200_APERM")) cfg->aperm = atoi(e) != 0 ? 1 : 0;
    cfg->pdl = batch_mode ? 0 : 1;       // every grid of the iteration fits the GPU at once: early launches cannot starve the running kernel
    if (const char *e = getenv("ICP_B200_PDL")) cfg->pdl = atoi(e) != 0 ? 1 : 0;
    cfg->fuseD = batch_mode ? 1 : 0;
    if (const char *e = getenv("ICP_B200_FUSED")) cfg->fuseD = atoi(e) != 0 ? 1 : 0;
    cfg->GB = batch_mode ? 1u : 8u;
    cfg->TC = batch_mode ? (uint32_t)SORTED_WARPS * 32u : 256u;
    if (const char *e = getenv("ICP_B200_GB")) { int v = atoi(e); if (v >= 1 && v <= 64) cfg->GB = (uint32_t)v; }
    if (const char *e = getenv("ICP_B200_TC")) { int v = atoi(e); if (v >= 64 && v <= SORTED_WARPS * 32 && v % 32 == 0) cfg->TC = (uint32_t)v; }
    if (cfg->Cmode < 2 && cfg->QG == 2048u && batch_mode) cfg->QG = 1024u;
    if (const char *e = getenv("ICP_B200_QG")) { int v = atoi(e); if (v >= 32 && v <= 2048 && v % 4 == 0) cfg->QG = (uint32_t)v; }
    if (const char *e = getenv("ICP_B200_QI")) { int v = atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32) cfg->QI = (uint32_t)v; }
    cfg->item_ovh = 24u;      // measured flat between 24 and never-split (0.317-0.318 ms), worse below 12: the fixed cost of an item is high
    if (const char *e = getenv("ICP_B200_ITEM_OVH")) { int v = atoi(e); if (v >= 0 && v <= 100000) cfg->item_ovh = (uint32_t)v; }
    cfg->span_pts = 0u;
    if (cfg->Cmode == 3)
    {
        // geometry of C'': 512-thread CTAs own 1024 sorted positions (two CTAs per SM: one CTA's record loads / barriers overlap
        // the other's list scans), 1024-thread CTAs own 2048 (ICP_B200_TC=1024)
        cfg->TC = 512u;
        if (const char *e = getenv("ICP_B200_TC")) { int v = atoi(e); if (v == 1024) cfg->TC = 1024u; }
        cfg->QG = cfg->TC * 2u;
        if (batch_mode) cfg->GB = 4u;        // B'' moves 48 bytes per query: four CTAs per pair
        if (const char *e = getenv("ICP_B200_GB")) { int v = atoi(e); if (v >= 1 && v <= 64) cfg->GB = (uint32_t)v; }
        // window of the list span: what is left of the 227 KB of dynamic shared memory (4 KB spared for the static arrays of the
        // kernel and of the fused kernel-D tail), at most 4096 points, at least 1024 -- else the sorted flavour without the window
        const size_t fixed = span_carve(nullptr, nullptr, cfg->nr, cfg->QG, cfg->QI, 0u) + 4096u;
        const size_t cap = (cfg->TC == 512u) ? 113u * 1024u : 227u * 1024u;
        uint32_t pts = fixed < cap ? (uint32_t)((cap - fixed) / 32u) : 0u;
        if (pts > 4096u) pts = 4096u;
        pts &= ~63u;
        if (const char *e = getenv("ICP_B200_SPAN_PTS")) { int v = atoi(e); if (v >= 64 && (uint32_t)v <= pts) pts = (uint32_t)v & ~63u; }
        if (pts < 512u && !getenv("ICP_B200_SPAN_PTS")) { cfg->Cmode = 2; cfg->QG = 2048u; cfg->TC = (uint32_t)SORTED_WARPS * 32u; cfg->GB = batch_mode ? 1u : 8u; }
        else cfg->span_pts = pts;
    }
    if (cfg->Cmode == 2 && sorted_carve(nullptr, nullptr, cfg->nr, cfg->QG, cfg->QI) > 200u * 1024u) cfg->Cmode = 1;
    if (cfg->Cmode == 1 && (cfg->nr > 65535u || grouped_smem_bytes(*cfg) > 200u * 1024u)) cfg->Cmode = 0;
    // stage-2 pruned walk inside kernel A: needs the pruned kernel A and the grouped kernel C (which finishes the matched queries)
    // Measured (B200, 256 pairs): the walk settles 40-90 % of the queries and halves kernel C, but its dependent gathers
    // (anchor, row, candidates) cost kernel A as much => opt-in (ICP_B200_NNWALK=1) until the walk is software-pipelined.
    cfg->nn_walk = 0;
    if (const char *e = getenv("ICP_B200_NNWALK")) { if (atoi(e) != 0 && cfg->Amode == 1 && cfg->Cmode >= 1) { cfg->nn_walk = 1; cfg->settle = 0; if (cfg->Cmode >= 2) { cfg->Cmode = 1; if (cfg->QG > 1024u) cfg->QG = 1024u; } } }
    // kernel D's generic path (one 512-thread CTA per pair in the batch engine, a 16-CTA cluster for one large registration): cp.async
    // rings for phases 2 / 3.  Whole level-1 blocks only, and the kernel whose tail runs D must own enough dynamic shared memory
    // (the stand-alone kernel D is launched with it).
    {
        const size_t need = (22u * D_SSTRIDE + 2u * 11u * 128u) * sizeof(float) + DRING_BYTES;
        bool ok = cfg->CL == 1 && (m % 2048u) == 0u;     // (a 16-CTA cluster gains nothing from the rings: 320 vs 322 us at 307200 / 1024 -- 16 SMs are the limit there)
        if (cfg->Cmode == 2 && sorted_carve(nullptr, nullptr, cfg->nr, cfg->QG, cfg->QI) < need) ok = false;
        if (cfg->Cmode == 3 && span_carve(nullptr, nullptr, cfg->nr, cfg->QG, cfg->QI, cfg->span_pts) < need - DRING_BYTES + DRING_BYTES_T(cfg->TC)) ok = false;
        cfg->dring = ok ? 1 : 0;
        if (const char *e = getenv("ICP_B200_DRING")) { if (atoi(e) == 0) cfg->dring = 0; }
    }
}

static size_t assign_smem(const FusedCfg &cfg)
{
    return assign_smem_bytes(cfg.nr, cfg.QB, cfg.par_rank) + (cfg.Amode == 1 ? (size_t)cfg.QB * 4 + 16 : 0);
}
static size_t reduce_smem(int CL, bool ring = false, uint32_t T = 512u)
{
    size_t n = 22u * D_SSTRIDE + 2u * 11u * 128u;
    if (CL >= 8) n += 7u * 2048u + 128u + 6u * 128u + 11u * 8u;      // fast path: the CTA's points + exchanged partials
    const size_t with_ring = (22u * D_SSTRIDE + 2u * 11u * 128u) * sizeof(float) + DRING_BYTES_T(T);
    n *= sizeof(float);
    return (ring && with_ring > n) ? with_ring : n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember what was set on WHICH device
// (a process may hold contexts on several GPUs).  `seen` = per-call-site table, one slot per device ordinal.
#define ICP_MAX_DEVICES 64
template <typename K>
static int ensure_dyn_smem(K kern, size_t smem, size_t (&seen)[ICP_MAX_DEVICES], bool always = false)
{
    int dev = 0;
    ICP_CUDA(cudaGetDevice(&dev));
    size_t &slot = seen[(unsigned)dev % ICP_MAX_DEVICES];
    if ((always || smem > 48 * 1024) && smem + 1 > slot)
    {
        ICP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        slot = smem + 1;
    }
    return ICP_OK;
}

// cudaLaunchKernelEx with the optional cluster / programmatic-dependent-launch attributes
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, unsigned cluster, Args... args)
{
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1u)
    {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = cluster; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl)
    {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    lc.attrs = attr; lc.numAttrs = n;
    return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);
}

template <int S, int QPT, bool SEARCH>
static int launch_assign_sq(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs)
{
    const size_t smem = assign_smem(cfg);
    static size_t seen[ICP_MAX_DEVICES];
    ICP_CHECK(ensure_dyn_smem(k_assign<S, QPT, SEARCH>, smem, seen));
    k_assign<S, QPT, SEARCH><<<dim3(cfg.nbA, n_pairs), cfg.TPB, smem, st>>>(table, cfg);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

template <int S, bool SEARCH>
static int launch_assign_s(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs)
{
    if (cfg.QPT == 4) return launch_assign_sq<S, 4, SEARCH>(st, cfg, table, n_pairs);
    return launch_assign_sq<S, 2, SEARCH>(st, cfg, table, n_pairs);
}

// the triangle bound of k_assign_tri is proven for metric weights in [0, 1] (any alpha >= 0)
static inline int tri_metric_ok(const FusedCfg &cfg) { return cfg.fg >= 0.f && cfg.fg <= 1.f && cfg.fp >= 0.f && cfg.fp <= 1.f; }

template <bool SEARCH>
static int launch_assign(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, bool pdl = false)
{
    if (cfg.Amode == 1)
    {
        const size_t smem = assign_smem(cfg);
        static size_t seen[ICP_MAX_DEVICES], seen_p[ICP_MAX_DEVICES], seen_s[ICP_MAX_DEVICES], seen_ps[ICP_MAX_DEVICES];
        const dim3 grid(cfg.nbA, n_pairs), block(cfg.TPB);
        const bool settle = SEARCH && cfg.settle != 0 && cfg.nn_walk == 0;
        if (SEARCH && cfg.aperm && settle)
        {
            ICP_CHECK(ensure_dyn_smem(k_assign_tri<SEARCH, SEARCH, SEARCH>, smem, seen_ps));
            ICP_CUDA(launch_k(k_assign_tri<SEARCH, SEARCH, SEARCH>, grid, block, smem, st, SEARCH && pdl, 1u, table, cfg, tri_metric_ok(cfg)));
        }
        else if (SEARCH && cfg.aperm)
        {
            ICP_CHECK(ensure_dyn_smem(k_assign_tri<SEARCH, SEARCH, false>, smem, seen_p));
            ICP_CUDA(launch_k(k_assign_tri<SEARCH, SEARCH, false>, grid, block, smem, st, SEARCH && pdl, 1u, table, cfg, tri_metric_ok(cfg)));
        }
        else if (settle)
        {
            ICP_CHECK(ensure_dyn_smem(k_assign_tri<SEARCH, false, SEARCH>, smem, seen_s));
            ICP_CUDA(launch_k(k_assign_tri<SEARCH, false, SEARCH>, grid, block, smem, st, SEARCH && pdl, 1u, table, cfg, tri_metric_ok(cfg)));
        }
        else
        {
            ICP_CHECK(ensure_dyn_smem(k_assign_tri<SEARCH, false, false>, smem, seen));
            ICP_CUDA(launch_k(k_assign_tri<SEARCH, false, false>, grid, block, smem, st, SEARCH && pdl, 1u, table, cfg, tri_metric_ok(cfg)));
        }
        return ICP_OK;
    }
    switch (cfg.S)
    {
        case 1: return launch_assign_s<1, SEARCH>(st, cfg, table, n_pairs);
        case 2: return launch_assign_s<2, SEARCH>(st, cfg, table, n_pairs);
        case 4: return launch_assign_s<4, SEARCH>(st, cfg, table, n_pairs);
        case 8: return launch_assign_s<8, SEARCH>(st, cfg, table, n_pairs);
        case 16: return launch_assign_s<16, SEARCH>(st, cfg, table, n_pairs);
        default: return launch_assign_s<32, SEARCH>(st, cfg, table, n_pairs);
    }
}

template <int CL, int T>
__global__ void __launch_bounds__(T, 1024 / T) k_reduce_solve(const PairPtrs *__restrict__ table, const FusedCfg cfg,
                                                        cudaGraphConditionalHandle handle, int use_handle)
{
    extern __shared__ float smem_d_k[];
    pdl_wait(); pdl_trigger();
    const PairPtrs P = table[blockIdx.y];
    reduce_solve_body<CL, T>(P, cfg, handle, use_handle, smem_d_k, (CL == 1) ? 0u : blockIdx.x);
}

// Wide flavour of kernel D for ONE large registration (m >= 65536): a 16-CTA cluster keeps 16 of the 148 SMs busy with 8.6 MB of
// sorted rows per pass (307200 points: 47.7 of 311 us per iteration, tools/scaled_phases.py).  Here every pass is an ordinary
// launch over the whole GPU (the level-1 blocks strided over gridDim.x CTAs, same slots / trees / order => same bits) and the
// kernel boundary replaces the cluster barrier: 3 wide launches + 1 single-CTA launch (second levels, solve, pose update).
template <int T>
__global__ void __launch_bounds__(T, 1024 / T) k_reduce_wide(const PairPtrs *__restrict__ table, const FusedCfg cfg,
                                                       cudaGraphConditionalHandle handle, int use_handle, int phase)
{
    extern __shared__ float smem_d_w[];
    pdl_wait(); pdl_trigger();
    const PairPtrs P = table[blockIdx.y];
    reduce_solve_body<0, T>(P, cfg, handle, use_handle, smem_d_w, blockIdx.x, phase);
}

// =================================================================================================
// Persistent cooperative iteration engine (latency mode; SURVEY 7 step 5(ii), north_star: "a persistent cooperative-grid or
// CUDA-Graph-captured iteration ... pick by measurement").  ONE cooperative launch runs n iterations of one registration:
// the four phases are the bodies of kernels A, B, C (CTAs loop over the virtual blocks) and D (the first 8-CTA cluster of
// the grid: same distributed-shared-memory path as k_reduce_solve<8, .>), separated by a software grid barrier instead of a
// kernel boundary; ICP::check () / the iteration budget are read from device memory at the top of every trip.  Same device
// functions, same arithmetic => bit-identical to the graph engine.  Everything another CTA produced inside the launch is
// read with ld.global.cg / acquire loads (L1 is not coherent): the bodies already do so for PDL's sake, and the pose T too.
// =================================================================================================
// HIER: the grid is made of 8-CTA clusters -- the CTAs of a cluster meet at the hardware cluster barrier and only one thread
// per cluster takes part in the global arrival count (16 serialised atomics on one L2 line instead of 128).
template <bool HIER>
__device__ __forceinline__ void grid_barrier(uint32_t *ctr, uint32_t &target, const uint32_t nblocks)
{
    if (HIER)
    {
        cg::cluster_group cluster = cg::this_cluster();
        target += nblocks / 8u;
        cluster.sync();
        if (cluster.block_rank() == 0 && threadIdx.x == 0)
        {
            __threadfence();
            atomicAdd(ctr, 1u);
            uint32_t v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
            __threadfence();
        }
        cluster.sync();
        return;
    }
    target += nblocks;                           // monotonic arrival count: no sense reversal, no reset inside the launch
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();                         // this CTA's writes (ordered before by the barrier above) are visible device-wide
        atomicAdd(ctr, 1u);
        uint32_t v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory"); } while (v < target);
    }
    __syncthreads();
}

// Phase D is a separate (non-inlined) function: it gets its own register allocation and schedule, as in its own kernel.
// Measured on B200 (tools/latency_engines.py): inlined into one body with A / B / C, the single-warp power method picked up
// spills (+4 us per iteration); with ALL phases non-inlined the pair / configuration structs live in local memory and A / B / C
// lose 2.4 us.  So: A, B, C inlined, D not.
template <int T>
__device__ __noinline__ void persist_phase_D(const PairPtrs &P, const FusedCfg &cfg, float *smem, uint32_t rank) { reduce_solve_body<8, T>(P, cfg, 0, 0, smem, rank); }

template <int T, bool HIER>
__global__ void __launch_bounds__(T, 1) k_icp_persistent(const PairPtrs *__restrict__ table, const FusedCfg cfg, const int tri_cfg, const uint32_t n_iters)
{
    extern __shared__ __align__(16) float4 smem_p[];
    const PairPtrs P = table[0];
    const uint32_t nb = gridDim.x;
    const uint32_t nbA = cfg.nbA, nbB = (cfg.nr + 31u) / 32u, nbC = (cfg.m + cfg.QG - 1u) / cfg.QG;
    uint32_t target = 0u;
    unsigned long long t0 = 0;
    if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) t0 = gtime_ns();
    for (uint32_t it = 0; it < n_iters; ++it)
    {
        if (__ldcg(&P.state->done)) break;       // written by phase D of the previous trip, before the last barrier: uniform
        // the launcher sizes the chunks so that every phase has at most one virtual block per CTA: the block index stays the
        // hardware blockIdx.x (the compiler keeps its uniform-datapath address arithmetic, as in the stand-alone kernels)
        if (blockIdx.x < nbA) assign_tri_body<true, false, false>(P, cfg, tri_cfg, blockIdx.x, smem_p);
        grid_barrier<HIER>(P.gbar, target, nb);
        if (blockIdx.x < nbB) colscan_body<true>(P, cfg, blockIdx.x);
        grid_barrier<HIER>(P.gbar, target, nb);
        if (blockIdx.x < nbC) search_grouped_body<false, T / 32>(P, cfg, blockIdx.x, smem_p);
        grid_barrier<HIER>(P.gbar, target, nb);
        if (blockIdx.x < 8u) persist_phase_D<T>(P, cfg, reinterpret_cast<float *>(smem_p), blockIdx.x);
        grid_barrier<HIER>(P.gbar, target, nb);
    }
    if (P.prof && blockIdx.x == 0 && threadIdx.x == 0) P.prof[60] = gtime_ns() - t0;
}

__global__ void k_fused_reps(const PairPtrs *__restrict__ table, uint32_t m, uint32_t W, uint32_t nrx, uint32_t nry, uint32_t sx, uint32_t sy)
{
    const PairPtrs P = table[blockIdx.y];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0)
    {
        P.wconst[0] = 1u;               // cleared by kernel A (build) if a fixed point breaks the constant-w property
        P.wconst[1] = 1u;               // cleared by k_rep_neighbours if a representative distance is not finite
        P.wconst[2] = 0u;               // arrival counter of k_search_sorted<true> (fused kernel-D tail)
        P.wconst[12] = 0u;              // no lane order from a previous iteration yet
        P.wconst[13] = 1u;              // k_build_scatter resets every temporal bound: the (empty) bound state is consistent
    }
    // guess of every fixed point's representative (seed of the build pass): the cell of the sampling grid it lies in
    if (t < m)
    {
        const uint32_t x = t % W, y = t / W;
        P.rep_id[t] = min(y / sy, nry - 1u) * nrx + min(x / sx, nrx - 1u);
    }
    if (t >= nrx * nry * 2u) return;
    const uint32_t r = t >> 1, h = t & 1u;
    const uint32_t gy = r / nrx, gx = r % nrx;
    const uint32_t xi = gx * sx + (sx >> 1) - 1u, yi = gy * sy + (sy >> 1) - 1u;
    ((float4 *)P.reps)[t] = __ldg((const float4 *)P.F + ((size_t)yi * W + xi) * 2u + h);
}

// ICPStep::buildRBC for every pair of the table
int fused_launch_build(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, uint32_t lm_w, uint32_t lm_h)
{
    uint32_t nrx, nry;
    icp_rep_grid(cfg.nr, &nrx, &nry);
    k_fused_reps<<<dim3(div_up(cfg.m > cfg.nr * 2 ? cfg.m : cfg.nr * 2, 256), n_pairs), 256, 0, st>>>(table, cfg.m, lm_w, nrx, nry, lm_w / nrx, lm_h / nry);
    ICP_LAUNCH_CHECK();
    if (cfg.Amode == 1)
    {
        k_rep_neighbours<<<dim3(cfg.nr, n_pairs), 256, (size_t)cfg.nr * 4, st>>>(table, cfg);
        ICP_LAUNCH_CHECK();
    }
    ICP_CHECK(launch_assign<false>(st, cfg, table, n_pairs));
    k_colscan<false><<<dim3(div_up(cfg.nr, 32), n_pairs), 1024, 0, st>>>(table, cfg);
    ICP_LAUNCH_CHECK();
    k_build_scatter<<<dim3(div_up(cfg.m, 256), n_pairs), 256, (size_t)cfg.nr * 4, st>>>(table, cfg);
    ICP_LAUNCH_CHECK();
    if (cfg.nn_walk)
    {
        k_list_neighbours<<<dim3(div_up(cfg.m, 128), n_pairs), 128, 0, st>>>(table, cfg);
        ICP_LAUNCH_CHECK();
    }
    return ICP_OK;
}

template <int CL, int T>
static int launch_reduce_solve(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs,
                               cudaGraphConditionalHandle handle, int use_handle, bool pdl)
{
    const size_t smem = reduce_smem(CL, cfg.dring != 0 && (T % 512 == 0), (uint32_t)T);
    static size_t seen[ICP_MAX_DEVICES];
    ICP_CHECK(ensure_dyn_smem(k_reduce_solve<CL, T>, smem, seen, true));
    ICP_CUDA(launch_k(k_reduce_solve<CL, T>, dim3(CL, n_pairs, 1), dim3(T, 1, 1), smem, st, pdl, (unsigned)CL, table, cfg, handle, use_handle));
    return ICP_OK;
}

static int launch_search(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, bool fuse_d, bool pdl = false);
static bool fuse_d_ok(const FusedCfg &cfg);
static int launch_colscan(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, bool pdl = false);

// latency mode: one 8-CTA cluster of 1024 threads per pair; batch mode: one CTA per pair, 256 threads by default so that
// several pairs share an SM and hide each other's dependent phases (cfg.TD)
static int launch_reduce_solve_cfg(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs,
                                   cudaGraphConditionalHandle handle, int use_handle, bool pdl = false)
{
    if (cfg.wideD)
    {
        const size_t smem = reduce_smem(1);
        static size_t seen[ICP_MAX_DEVICES];
        ICP_CHECK(ensure_dyn_smem(k_reduce_wide<1024>, smem, seen, true));
        const uint32_t nb128p = (div_up(cfg.m, 128u) + 3u) & ~3u, nb512 = div_up(div_up(cfg.m, 4u), 512u);
        uint32_t G = div_up(nb128p, 32u);                   // one level-1 block per warp ...
        if (div_up(nb512, 2u) > G) G = div_up(nb512, 2u);   // ... two S blocks per CTA and round
        if (G > 128u) G = 128u;
        for (int phase = 1; phase <= 4; ++phase)
            ICP_CUDA(launch_k(k_reduce_wide<1024>, dim3(phase == 4 ? 1u : G, n_pairs, 1), dim3(1024, 1, 1), smem, st, pdl, 1u, table, cfg, handle, use_handle, phase));
        return ICP_OK;
    }
    if (cfg.CL == 16)
    {
        // non-portable cluster size: needs the opt-in attribute and 16 free SMs in one GPC; 8 otherwise
        static int ok16[ICP_MAX_DEVICES];          // 0 = not tried, 1 = available, -1 = unavailable
        int dev = 0;
        cudaGetDevice(&dev);
        int &st16 = ok16[(unsigned)dev % ICP_MAX_DEVICES];
        if (st16 == 0)
        {
            // asked once per device, with occupancy queries only (this may run inside a stream capture: no trial launch)
            st16 = -1;
            const size_t smem = reduce_smem(16, cfg.dring != 0, 1024u);
            if (cudaFuncSetAttribute(k_reduce_solve<16, 1024>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                cudaFuncSetAttribute(k_reduce_solve<16, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess)
            {
                cudaLaunchConfig_t lc;
                memset(&lc, 0, sizeof(lc));
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = 16; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                lc.attrs = at; lc.numAttrs = 1; lc.gridDim = dim3(16, 1, 1); lc.blockDim = dim3(1024, 1, 1); lc.dynamicSmemBytes = smem;
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, k_reduce_solve<16, 1024>, &lc) == cudaSuccess && nc >= 1) st16 = 1;
            }
            cudaGetLastError();
        }
        if (st16 == 1) return launch_reduce_solve<16, 1024>(st, cfg, table, n_pairs, handle, use_handle, pdl);
        return launch_reduce_solve<8, 1024>(st, cfg, table, n_pairs, handle, use_handle, pdl);
    }
    if (cfg.CL == 8) return launch_reduce_solve<8, 1024>(st, cfg, table, n_pairs, handle, use_handle, pdl);
    if (cfg.TD == 256) return launch_reduce_solve<1, 256>(st, cfg, table, n_pairs, handle, use_handle, pdl);
    if (cfg.TD == 512) return launch_reduce_solve<1, 512>(st, cfg, table, n_pairs, handle, use_handle, pdl);
    return launch_reduce_solve<1, 1024>(st, cfg, table, n_pairs, handle, use_handle, pdl);
}

int fused_launch_one(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, int which)
{
    switch (which)
    {
        case 0: return launch_assign<true>(st, cfg, table, n_pairs);
        case 1: return launch_colscan(st, cfg, table, n_pairs);
        case 2: return launch_search(st, cfg, table, n_pairs, false);
        default:
            return launch_reduce_solve_cfg(st, cfg, table, n_pairs, 0, 0);
    }
}

int fused_launch_iteration(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs,
                           cudaGraphConditionalHandle handle, int use_handle)
{
    // latency mode: programmatic dependent launch along the chain A -> B -> C -> D -> A (grouped kernel C only)
    // Measured (B200, one pair): plain stream launches 52.8 -> 45.4 us per iteration; inside a captured graph the node-to-node
    // gap is already ~0.5 us and what remains between two kernels is the completion + flush of the predecessor, which
    // the programmatic edge does not remove (44.2 vs 43.8 us) => only used when the stream is not being captured.
    bool pdl = cfg.pdl != 0 && !use_handle && cfg.Amode == 1 && cfg.Cmode == 1;
    if (pdl)
    {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) pdl = false;
    }
    ICP_CHECK(launch_assign<true>(st, cfg, table, n_pairs, pdl));
    ICP_CHECK(launch_colscan(st, cfg, table, n_pairs, pdl));
    if (!use_handle && fuse_d_ok(cfg)) return launch_search(st, cfg, table, n_pairs, true);      // kernel D runs in the tail of C'
    ICP_CHECK(launch_search(st, cfg, table, n_pairs, false, pdl));
    return launch_reduce_solve_cfg(st, cfg, table, n_pairs, handle, use_handle, pdl);
}

static size_t grouped_smem_bytes(const FusedCfg &cfg) { return grouped_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI); }
static size_t grouped_smem(const FusedCfg &cfg) { return grouped_smem_bytes(cfg); }

static int launch_colscan(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, bool pdl)
{
    if (cfg.Cmode == 2)
    {
        const size_t smem = colsort_smem_bytes(cfg);
        static size_t seen_m[ICP_MAX_DEVICES], seen_s[ICP_MAX_DEVICES];
        const int multi = cfg.GB > 1u ? 1 : 0;
        if (multi) ICP_CHECK(ensure_dyn_smem(k_colscan_sort<true, false>, smem, seen_m));
        else ICP_CHECK(ensure_dyn_smem(k_colscan_sort<false, false>, smem, seen_s));
        if (multi) k_colscan_sort<true, false><<<dim3(cfg.GB, n_pairs), COLSORT_THREADS, smem, st>>>(table, cfg);
        else k_colscan_sort<false, false><<<dim3(1, n_pairs), COLSORT_THREADS, smem, st>>>(table, cfg);
    }
    else if (cfg.Cmode == 3)
    {
        const size_t smem = colsort_smem_bytes(cfg);
        static size_t seen_m[ICP_MAX_DEVICES], seen_s[ICP_MAX_DEVICES];
        const int multi = cfg.GB > 1u ? 1 : 0;
        if (multi) ICP_CHECK(ensure_dyn_smem(k_colscan_sort<true, true>, smem, seen_m));
        else ICP_CHECK(ensure_dyn_smem(k_colscan_sort<false, true>, smem, seen_s));
        if (multi) k_colscan_sort<true, true><<<dim3(cfg.GB, n_pairs), COLSORT_THREADS, smem, st>>>(table, cfg);
        else k_colscan_sort<false, true><<<dim3(1, n_pairs), COLSORT_THREADS, smem, st>>>(table, cfg);
    }
    else ICP_CUDA(launch_k(k_colscan<true>, dim3(div_up(cfg.nr, 32), n_pairs), dim3(1024), 0, st, pdl, 1u, table, cfg));
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

static bool fuse_d_ok(const FusedCfg &cfg)
{
    if (cfg.Cmode == 3) return cfg.fuseD && cfg.CL == 1 && span_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI, cfg.span_pts) >= reduce_smem(1);
    return cfg.fuseD && cfg.Cmode == 2 && cfg.CL == 1 && cfg.TC == (uint32_t)SORTED_WARPS * 32u
           && sorted_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI) >= reduce_smem(1);
}

static int launch_search(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, bool fuse_d, bool pdl)
{
    if (cfg.Cmode == 3)
    {
        const size_t smem = span_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI, cfg.span_pts);
        static size_t seen_f[ICP_MAX_DEVICES], seen_u[ICP_MAX_DEVICES];
        static size_t seen_f5[ICP_MAX_DEVICES], seen_u5[ICP_MAX_DEVICES];
        const dim3 grid(div_up(cfg.m, cfg.QG), n_pairs);
        if (cfg.TC == 512u)
        {
            if (fuse_d) ICP_CHECK(ensure_dyn_smem(k_search_span<true, 512>, smem, seen_f5, true));
            else ICP_CHECK(ensure_dyn_smem(k_search_span<false, 512>, smem, seen_u5, true));
            if (fuse_d) k_search_span<true, 512><<<grid, 512, smem, st>>>(table, cfg);
            else k_search_span<false, 512><<<grid, 512, smem, st>>>(table, cfg);
        }
        else
        {
            if (fuse_d) ICP_CHECK(ensure_dyn_smem(k_search_span<true, 1024>, smem, seen_f, true));
            else ICP_CHECK(ensure_dyn_smem(k_search_span<false, 1024>, smem, seen_u, true));
            if (fuse_d) k_search_span<true, 1024><<<grid, 1024, smem, st>>>(table, cfg);
            else k_search_span<false, 1024><<<grid, 1024, smem, st>>>(table, cfg);
        }
        ICP_LAUNCH_CHECK();
        return ICP_OK;
    }
    if (cfg.Cmode == 2)
    {
        const size_t smem = sorted_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI);
        static size_t seen_f[ICP_MAX_DEVICES], seen_u[ICP_MAX_DEVICES];
        const dim3 grid(div_up(cfg.m, cfg.QG), n_pairs);
        if (fuse_d) ICP_CHECK(ensure_dyn_smem(k_search_sorted<true>, smem, seen_f));
        else ICP_CHECK(ensure_dyn_smem(k_search_sorted<false>, smem, seen_u));
        if (fuse_d) k_search_sorted<true><<<grid, cfg.TC, smem, st>>>(table, cfg);
        else k_search_sorted<false><<<grid, cfg.TC, smem, st>>>(table, cfg);
        ICP_LAUNCH_CHECK();
        return ICP_OK;
    }
    if (cfg.Cmode == 1)
    {
        const size_t smem = grouped_smem(cfg);
        static size_t seen[ICP_MAX_DEVICES], seen_s[ICP_MAX_DEVICES];
        if (cfg.settle && !cfg.nn_walk)
        {
            ICP_CHECK(ensure_dyn_smem(k_search_grouped<true>, smem, seen_s));
            ICP_CUDA(launch_k(k_search_grouped<true>, dim3(div_up(cfg.m, cfg.QG), n_pairs), dim3(GROUPED_WARPS * 32), smem, st, pdl, 1u, table, cfg));
            return ICP_OK;
        }
        ICP_CHECK(ensure_dyn_smem(k_search_grouped<false>, smem, seen));
        ICP_CUDA(launch_k(k_search_grouped<false>, dim3(div_up(cfg.m, cfg.QG), n_pairs), dim3(GROUPED_WARPS * 32), smem, st, pdl, 1u, table, cfg));
        return ICP_OK;
    }
    const dim3 grid(div_up(cfg.m, cfg.QC), n_pairs);
    const size_t smem = (size_t)cfg.nr * 4;
    switch (cfg.L)
    {
        case 1: k_search<1><<<grid, 256, smem, st>>>(table, cfg); break;
        case 2: k_search<2><<<grid, 256, smem, st>>>(table, cfg); break;
        case 4: k_search<4><<<grid, 256, smem, st>>>(table, cfg); break;
        case 16: k_search<16><<<grid, 256, smem, st>>>(table, cfg); break;
        case 32: k_search<32><<<grid, 256, smem, st>>>(table, cfg); break;
        default: k_search<8><<<grid, 256, smem, st>>>(table, cfg); break;
    }
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

int fused_launch_one(cudaStream_t st, const FusedCfg &cfg, const PairPtrs *table, uint32_t n_pairs, int which);

// ------------------------------------------------------------------------------------------------
// single-pair engine glue
// ------------------------------------------------------------------------------------------------
struct FusedWS
{
    PairPtrs *table;     // device, 1 entry
    uint16_t *lrank;
    uint32_t *H;
    float *fxyz, *mxyz, *red;
    unsigned long long *prof;
    uint32_t *wconst;
    uint2 *nbr;
    uint32_t *nbx, *nn_o;
    float *nnd;
    float *Qs;
    uint4 *Rs;
    uint32_t *gbar;
    uint2 *QR, *nn2;
};

static size_t fused_ws_layout(uint32_t m, uint32_t nr, int sm_count, void *base, FusedWS *ws)
{
    Carver cv(base);
    // worst case rows of H: QB >= 32
    FusedCfg cfg;
    fused_choose_cfg(&cfg, m, nr, sm_count, 1);
    PairPtrs *table = cv.take<PairPtrs>(1);
    uint16_t *lrank = cv.take<uint16_t>((size_t)m + 8);
    uint32_t *H = cv.take<uint32_t>((size_t)cfg.nbA * nr + 32);
    float *fxyz = cv.take<float>((size_t)3 * m);
    float *mxyz = cv.take<float>((size_t)3 * m);
    float *red = cv.take<float>(fused_red_elems(m));
    unsigned long long *prof = cv.take<unsigned long long>(64);
    uint32_t *wconst = cv.take<uint32_t>(16);
    uint2 *nbr = cv.take<uint2>(fused_nbr_elems(nr));
    uint32_t *nbx = cv.take<uint32_t>((size_t)m * FUSED_NBX_K + 8);
    uint32_t *nn_o = cv.take<uint32_t>(m);
    float *nnd = cv.take<float>(m);
    float *Qs = cv.take<float>((size_t)m * 8);
    uint4 *Rs = cv.take<uint4>(m);
    uint32_t *gbar = cv.take<uint32_t>(4);
    uint2 *QR = cv.take<uint2>(m);
    uint2 *nn2 = cv.take<uint2>(m);
    if (ws) { ws->nn2 = nn2; ws->QR = QR; ws->gbar = gbar; ws->Qs = Qs; ws->Rs = Rs; ws->nbx = nbx; ws->nn_o = nn_o; ws->nnd = nnd; ws->nbr = nbr; ws->wconst = wconst; ws->prof = prof; ws->table = table; ws->lrank = lrank; ws->H = H; ws->fxyz = fxyz; ws->mxyz = mxyz; ws->red = red; }
    return cv.off + 256;
}

size_t fused_workspace_bytes(uint32_t m, uint32_t nr, int sm_count) { return fused_ws_layout(m, nr, sm_count, nullptr, nullptr); }

int fused_prepare(icp_step *s)
{
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    PairPtrs P;
    memset(&P, 0, sizeof(P));
    P.F = s->F; P.M = s->M; P.T = s->T; P.reps = s->reps; P.Xp = s->Xp; P.N = s->N; P.O = s->O;
    P.rep_id = s->rep_id; P.perm = s->perm; P.q_rep = s->q_rep; P.lrank = ws.lrank; P.H = ws.H;
    P.Nq = s->Nq; P.Oq = s->Oq; P.qperm = s->qperm; P.W = s->W; P.fxyz = ws.fxyz; P.mxyz = ws.mxyz; P.NNID = s->NNID;
    P.sum_w = s->sum_w; P.mean = s->mean; P.S = s->S; P.Tk = s->Tk; P.Rk = s->Rk; P.state = s->state; P.loop = s->loop;
    P.evals = s->count_evals ? s->evals : nullptr;
    P.red = ws.red;
    P.prof = ws.prof;
    P.wconst = ws.wconst;
    P.nbr = ws.nbr;
    P.nbx = ws.nbx; P.nn_o = ws.nn_o; P.nnd = ws.nnd;
    P.Qs = ws.Qs; P.Rs = ws.Rs; P.gbar = ws.gbar; P.QR = ws.QR; P.nn2 = ws.nn2;
    // tiny, rare: synchronous upload keeps the table consistent with the graphs captured afterwards
    ICP_CUDA(cudaStreamSynchronize(s->ctx->stream));
    ICP_CUDA(cudaMemcpy(ws.table, &P, sizeof(P), cudaMemcpyHostToDevice));
    return ICP_OK;
}

static void fused_cfg_of(icp_step *s, FusedCfg *cfg)
{
    fused_choose_cfg(cfg, s->m, s->nr, s->ctx->sm_count, 1);
    cfg->lm_w = s->lm_w; cfg->lm_h = s->lm_h;
    // grouped kernel C over 32 x 16 patches of the landmark grid (see search_grouped_body) for one large image-ordered registration
    cfg->ctile_w = 0u;
    // 32 x 24 patches measured best at 640 x 480 (331 / 269 us per iteration at 512 / 1024 representatives against 339 / 277 with 32 x 16;
    // patches aligned with the 20 x 15 representative cells -- 40 x 15, 40 x 30, 20 x 30 -- 334-342 / 277-285)
    if (cfg->Cmode == 1 && s->m >= 131072u && cfg->QG == 512u && s->lm_w % 32u == 0u && s->lm_h % 24u == 0u && !getenv("ICP_B200_QG")) cfg->QG = 768u;
    if (cfg->Cmode == 1 && s->m >= 65536u && cfg->QG >= 128u && cfg->QG % 32u == 0u && s->lm_w % 32u == 0u && s->lm_h % (cfg->QG / 32u) == 0u
        && (uint64_t)s->lm_w * s->lm_h == s->m) cfg->ctile_w = 32u;           // 32 x (QG / 32) patches
    if (const char *e = getenv("ICP_B200_CTILE")) { if (atoi(e) == 0) cfg->ctile_w = 0u; }
    // experiment knob: any patch shape that tiles the grid (QG = w x h)
    if (const char *e = getenv("ICP_B200_CTILE_WH"))
    {
        unsigned w = 0, h = 0;
        if (sscanf(e, "%ux%u", &w, &h) == 2 && w >= 4u && h >= 1u && cfg->Cmode == 1 && s->lm_w % w == 0u && s->lm_h % h == 0u && w * h <= 2048u && (w * h) % 4u == 0u
            && (uint64_t)s->lm_w * s->lm_h == s->m)
        { cfg->ctile_w = w; cfg->QG = w * h; }
    }
    cfg->fg = s->fg; cfg->fp = s->fp; cfg->c = s->c;
    cfg->weighted = s->w_cfg; cfg->power_method = (s->rot_cfg == ICP_ROT_POWER_METHOD);
}

int fused_enqueue_build(icp_step *s, cudaStream_t st)
{
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    FusedCfg cfg;
    fused_cfg_of(s, &cfg);
    return fused_launch_build(st, cfg, ws.table, 1, s->lm_w, s->lm_h);
}

int fused_enqueue_iteration(icp_step *s, cudaStream_t st, cudaGraphConditionalHandle handle, int use_handle)
{
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    FusedCfg cfg;
    fused_cfg_of(s, &cfg);
    return fused_launch_iteration(st, cfg, ws.table, 1, handle, use_handle);
}

// Host-side invalidation of the fused engine's cross-iteration state (see PairPtrs::wconst [12], [13]):
//   lane_order: the next kernel A must not use lperm (the previous iteration was not a fused one);
//   bounds:     the temporal-pruning bounds may describe another moving set (start of every run call).
int fused_invalidate(icp_step *s, cudaStream_t st, bool lane_order, bool bounds)
{
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    if (lane_order && bounds) ICP_CUDA(cudaMemsetAsync(ws.wconst + 12, 0, 2 * sizeof(uint32_t), st));
    else if (lane_order) ICP_CUDA(cudaMemsetAsync(ws.wconst + 12, 0, sizeof(uint32_t), st));
    else if (bounds) ICP_CUDA(cudaMemsetAsync(ws.wconst + 13, 0, sizeof(uint32_t), st));
    return ICP_OK;
}

// ---- persistent cooperative engine -------------------------------------------------------------------------------
// Every phase runs in the same grid of T-thread CTAs (one per SM, launched as clusters of 8 for phase D): the chunk / run
// sizes follow from the number of co-resident CTAs.  *ok = 0 (and nothing enqueued) when the problem size is not eligible
// or the device cannot co-schedule the grid: the caller then uses the graph engine.
template <int T, bool HIER>
static int persistent_launch(icp_step *s, cudaStream_t st, uint32_t n_iters, int *ok)
{
    *ok = 0;
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    FusedCfg cfg;
    fused_cfg_of(s, &cfg);
    if (cfg.Amode != 1 || s->nr > 4096u) return ICP_OK;
    const uint32_t h_rows = cfg.nbA;                   // rows of H the workspace holds (fused_ws_layout)
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    lc.attrs = attr; lc.numAttrs = 2; lc.stream = st;
    lc.blockDim = dim3(T, 1, 1);
    static int max_clusters[ICP_MAX_DEVICES];          // per device; 0 = not asked yet, < 0 = unavailable
    static size_t smem_asked[ICP_MAX_DEVICES];
    uint32_t n_cta = (uint32_t)s->ctx->sm_count / 8u * 8u;
    const int dev = s->ctx->device % ICP_MAX_DEVICES;
    for (int pass = 0; pass < 2; ++pass)               // pass 0: geometry for all SMs -> occupancy query; pass 1: the real grid
    {
        uint32_t QB = (div_up(s->m, n_cta) + 3u) & ~3u;
        if (QB < 32u) QB = 32u;
        if (QB > 1024u) return ICP_OK;                 // chunks beyond the shared-memory ranking: graph engine only
        cfg.QB = QB; cfg.nbA = div_up(s->m, QB);
        cfg.QG = QB; cfg.QI = 8u; cfg.Cmode = 1; cfg.CL = 8; cfg.SF = 32; cfg.aperm = 0; cfg.settle = 0; cfg.nn_walk = 0; cfg.fuseD = 0; cfg.pdl = 0; cfg.dring = 0; cfg.wideD = 0; cfg.ctile_w = 0u;
        cfg.par_rank = (assign_smem_bytes(s->nr, QB, 1) <= 96u * 1024u) ? 1 : 0;
        size_t smem = assign_smem(cfg);
        const size_t sg = grouped_carve(nullptr, nullptr, cfg.nr, cfg.QG, cfg.QI, T / 32);
        if (sg > smem) smem = sg;
        if (reduce_smem(8) > smem) smem = reduce_smem(8);
        if (smem > 200u * 1024u) return ICP_OK;
        lc.dynamicSmemBytes = smem;
        if (pass == 0)
        {
            if (max_clusters[dev] == 0 || smem > smem_asked[dev])
            {
                if (cudaFuncSetAttribute(k_icp_persistent<T, HIER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return ICP_OK; }
                smem_asked[dev] = smem;
                lc.gridDim = dim3(n_cta, 1, 1);
                int nc = 0;
                if (cudaOccupancyMaxActiveClusters(&nc, k_icp_persistent<T, HIER>, &lc) != cudaSuccess || nc < 1) { cudaGetLastError(); max_clusters[dev] = -1; }
                else max_clusters[dev] = nc;
            }
            if (max_clusters[dev] < 1) return ICP_OK;
            if ((uint32_t)max_clusters[dev] * 8u < n_cta) n_cta = (uint32_t)max_clusters[dev] * 8u;
            if (const char *e = getenv("ICP_B200_PERSIST_CTAS")) { int v = atoi(e); if (v >= 8 && v % 8 == 0 && (uint32_t)v <= n_cta) n_cta = (uint32_t)v; }
        }
    }
    if (cfg.nbA > h_rows || cfg.nbA > n_cta || div_up(s->m, cfg.QG) > n_cta || div_up(s->nr, 32u) > n_cta) return ICP_OK;
    lc.gridDim = dim3(n_cta, 1, 1);
    ICP_CUDA(cudaMemsetAsync(ws.gbar, 0, 4 * sizeof(uint32_t), st));
    ICP_CUDA(cudaLaunchKernelEx(&lc, k_icp_persistent<T, HIER>, (const PairPtrs *)ws.table, cfg, tri_metric_ok(cfg), n_iters));
    *ok = 1;
    return ICP_OK;
}

int fused_enqueue_persistent(icp_step *s, cudaStream_t st, uint32_t n_iters, int *ok)
{
    int threads = 512;
    if (const char *e = getenv("ICP_B200_PERSIST_T")) { if (atoi(e) == 1024) threads = 1024; }
    // grid barrier flavour: flat arrival count (default) or hierarchical through the hardware cluster barriers -- measured
    // 1 us slower per barrier on B200 (two cluster barriers cost more than 112 fewer serialised atomics save)
    int hier = 0;
    if (const char *e = getenv("ICP_B200_PERSIST_HIER")) hier = atoi(e) != 0;
    if (threads == 1024) return hier ? persistent_launch<1024, true>(s, st, n_iters, ok) : persistent_launch<1024, false>(s, st, n_iters, ok);
    return hier ? persistent_launch<512, true>(s, st, n_iters, ok) : persistent_launch<512, false>(s, st, n_iters, ok);
}

void *fused_debug_ptr(icp_step *s, const char *name)
{
    FusedWS ws;
    fused_ws_layout(s->m, s->nr, s->ctx->sm_count, s->fused, &ws);
    if (!strcmp(name, "fxyz")) return ws.fxyz;
    if (!strcmp(name, "mxyz")) return ws.mxyz;
    if (!strcmp(name, "H")) return ws.H;
    if (!strcmp(name, "lrank")) return ws.lrank;
    if (!strcmp(name, "prof")) return ws.prof;
    return nullptr;
}
