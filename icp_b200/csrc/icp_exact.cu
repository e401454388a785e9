// icp_exact.cu -- EXACT nearest-neighbour search over the random ball cover (the variant of the RBC paper
// that the reference's README.md:4 leaves out: "frame-to-frame only"; needed for frame-to-model mapping, where the
// database is a model cloud and an approximate match is not good enough).
//
// The one-shot search of the ICP pipeline (RBCSearch as wired at algorithms.cpp:4520-4536) looks only at the list of the
// query's nearest representative.  The exact search must return what a brute-force scan of the whole database returns;
// the cover lets it skip most lists:
//     sqrt D is a metric (D = fg |dxyz1|^2 + fp |drgb1|^2 with fg, fp >= 0 is a squared Euclidean distance in a scaled
//     space), so for a representative r with list radius psi_r = max_{x in list r} sqrt D(x, r):
//         sqrt D(q, x) >= sqrt D(q, r) - psi_r          for every x of the list of r.
//     If that lower bound exceeds sqrt(best), no point of the list can beat or tie the best match found so far.
// Floating point: what is EVALUATED uses the reference arithmetic (dist8: individually rounded, no FMA) and the ordered
// update rule (smaller distance, then lower list position), so among the evaluated points the winner is the one a
// sequential strict-'<' scan of X_p from +inf finds.  Only the decision NOT to evaluate must be safe: computed distances
// carry a relative error below 1e-6 and an absolute one below 1e-36, the radius is rounded up, the bound
// down (directed rounding throughout), and the test keeps the safety factors of the temporal pruning:
//         lb = sqrt_rd (D~(q, r)) * (1 - 1e-6) - psi_r ;   skip  iff  lb > 0  and  lb^2 * (1 - 1e-5) > best + 1e-30 .
// NaN operands fail every comparison (the list is scanned); +inf bounds skip only against a finite best.
// Metric weights outside [0, 1] (icp_step_set_metric) are not covered by the error analysis: every list is scanned.
#include "icp_stages.cuh"
#include <math_constants.h>

// psi[r]: one warp per representative
__global__ void __launch_bounds__(256) k_list_radius(const float *__restrict__ Xp, const float *__restrict__ R, const uint32_t *__restrict__ O,
                                                     const uint32_t *__restrict__ N, uint32_t nr, float fg, float fp, float *__restrict__ psi)
{
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (r >= nr) return;
    const pt8 rep = ld_pt8(R, r);
    const uint32_t o = O[r], n = N[r];
    float mx = 0.f;
    bool bad = false;
    for (uint32_t k = o + lane; k < o + n; k += 32u)
    {
        const pt8 x = ld_pt8(Xp, k);
        const float d = dist8(x.lo, x.hi, rep.lo, rep.hi, fg, fp);
        bad = bad || (d != d);
        mx = fmaxf(mx, d);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, off));
    bad = __any_sync(FULL_MASK, bad);
    // true sqrt D <= sqrt (D~ / (1 - 1e-6)) <= sqrt_ru (D~) * (1 + 1e-6); a NaN distance makes the radius +inf: never skipped
    if (lane == 0) psi[r] = bad ? CUDART_INF_F : __fmul_ru(__fsqrt_ru(mx), 1.000001f);
}

__device__ __forceinline__ bool exact_skip(float dr, float psi, float best)
{
    const float lb = __fsub_rd(__fmul_rd(__fsqrt_rd(dr), 0.999999f), psi);
    return lb > 0.f && __fmul_rd(__fmul_rd(lb, lb), 0.99999f) > __fadd_ru(best, 1e-30f);
}

// all lanes of the warp scan list [o, o + n) for the same query; (best, bi) is warp-uniform on entry and on exit
__device__ __forceinline__ void exact_scan_list(const float *__restrict__ Xp, uint32_t o, uint32_t n, const pt8 &q, float fg, float fp,
                                                float &best, uint32_t &bi, uint32_t lane)
{
    float b = best;
    uint32_t id = bi;
#pragma unroll 2
    for (uint32_t k = o + lane; k < o + n; k += 32u)
    {
        const pt8 x = ld_pt8(Xp, k);
        const float d = dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);
        if (d < b || (d == b && k < id)) { b = d; id = k; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
    {
        const float od = __shfl_xor_sync(FULL_MASK, b, off);
        const uint32_t oi = __shfl_xor_sync(FULL_MASK, id, off);
        if (od < b || (od == b && oi < id)) { b = od; id = oi; }
    }
    best = b; bi = id;
}

#define EXACT_WARPS 8
// one warp per query; the representatives, list offsets / sizes and radii are staged in shared memory
__global__ void __launch_bounds__(EXACT_WARPS * 32) k_rbc_search_exact(const float *__restrict__ Q, uint32_t m, const float *__restrict__ R, uint32_t nr,
                                                                       const float *__restrict__ Xp, const uint32_t *__restrict__ O,
                                                                       const uint32_t *__restrict__ N, const float *__restrict__ psi, float fg, float fp,
                                                                       int prune, icp_dist_id *__restrict__ NNID, float *__restrict__ NN,
                                                                       unsigned long long *__restrict__ evals)
{
    extern __shared__ float4 smem_x[];
    float4 *sR = smem_x;                                         // [nr][2]
    uint32_t *sO = reinterpret_cast<uint32_t *>(sR + (size_t)nr * 2);
    uint32_t *sN = sO + nr;
    float *sPsi = reinterpret_cast<float *>(sN + nr);
    for (uint32_t i = threadIdx.x; i < nr * 2u; i += blockDim.x) sR[i] = __ldg((const float4 *)R + i);
    for (uint32_t r = threadIdx.x; r < nr; r += blockDim.x) { sO[r] = O[r]; sN[r] = N[r]; sPsi[r] = psi[r]; }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned long long ev = 0;
    for (uint32_t i = blockIdx.x * EXACT_WARPS + warp; i < m; i += gridDim.x * EXACT_WARPS)
    {
        const pt8 q = ld_pt8(Q, i);
        // nearest representative (ordered argmin): its list is scanned first, it gives a tight first bound
        float bd = CUDART_INF_F;
        uint32_t s = 0;
        for (uint32_t r = lane; r < nr; r += 32u)
        {
            const float d = dist8(q.lo, q.hi, sR[2 * r], sR[2 * r + 1], fg, fp);
            if (d < bd) { bd = d; s = r; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
        {
            const float od = __shfl_xor_sync(FULL_MASK, bd, off);
            const uint32_t oi = __shfl_xor_sync(FULL_MASK, s, off);
            if (od < bd || (od == bd && oi < s)) { bd = od; s = oi; }
        }
        float best = CUDART_INF_F;
        uint32_t bi = 0u;                        // the sequential scan from +inf keeps position 0 when nothing compares less
        exact_scan_list(Xp, sO[s], sN[s], q, fg, fp, best, bi, lane);
        ev += nr + sN[s];
        // every other list, 32 representatives at a time: the lanes test the bound, the warp scans the survivors one by one
        // (the bound tightens as best shrinks, so the remaining candidates are re-tested after every scan)
        for (uint32_t r0 = 0; r0 < nr; r0 += 32u)
        {
            const uint32_t r = r0 + lane;
            float dr = CUDART_INF_F, ps = 0.f;
            bool cand = r < nr && r != s && sN[r] > 0u;
            if (cand)
            {
                dr = dist8(q.lo, q.hi, sR[2 * r], sR[2 * r + 1], fg, fp);
                ps = sPsi[r];
                if (prune && exact_skip(dr, ps, best)) cand = false;
            }
            uint32_t mask = __ballot_sync(FULL_MASK, cand);
            while (mask)
            {
                const uint32_t l = (uint32_t)__ffs((int)mask) - 1u;
                mask &= mask - 1u;
                const uint32_t rr = r0 + l;
                exact_scan_list(Xp, sO[rr], sN[rr], q, fg, fp, best, bi, lane);
                ev += sN[rr];
                if (mask)
                {
                    if (cand && prune && exact_skip(dr, ps, best)) cand = false;
                    mask &= __ballot_sync(FULL_MASK, cand);
                }
            }
        }
        if (lane == 0)
        {
            icp_dist_id o2; o2.dist = best; o2.id = bi;
            NNID[i] = o2;
        }
        if (NN && lane < 2u) reinterpret_cast<float4 *>(NN)[(size_t)i * 2 + lane] = __ldg((const float4 *)Xp + (size_t)bi * 2 + lane);
    }
    if (evals && lane == 0 && ev) atomicAdd(evals, ev);
}

// RBC exact search entry (see include/icp_b200.h)
extern "C" int icp_rbc_search_exact(icp_ctx *ctx, const float *d_Q, uint32_t m, const float *d_R, uint32_t nr, float alpha,
                                    const float *d_Xp, const uint32_t *d_O, const uint32_t *d_N,
                                    icp_dist_id *d_NN_ID, float *d_NN, uint64_t *d_evals)
{
    ICP_ENTER(ctx);
    if (!ctx || !d_Q || !d_R || !d_Xp || !d_O || !d_N || !d_NN_ID) { icp_set_error("icp_rbc_search_exact: null argument"); return ICP_ERR_ARG; }
    if (m == 0) ICP_CONFIG_FAIL("RBCSearch", "The array Q cannot have zero points");
    if (nr == 0 || nr > 4096) ICP_CONFIG_FAIL("RBCSearch", "The array R must have between 1 and 4096 points");
    if (alpha == 0.f) ICP_CONFIG_FAIL("RBCSearch", "The alpha parameter cannot be equal to zero");
    float fg, fp;
    icp_metric_weights(alpha, &fg, &fp);
    const int prune = (fg >= 0.f && fg <= 1.f && fp >= 0.f && fp <= 1.f) ? 1 : 0;
    void *scr;
    ICP_CHECK(icp_ctx_scratch(ctx, (size_t)nr * sizeof(float) + 256, &scr));
    float *psi = (float *)scr;
    k_list_radius<<<div_up(nr * 32u, 256), 256, 0, ctx->stream>>>(d_Xp, d_R, d_O, d_N, nr, fg, fp, psi);
    ICP_LAUNCH_CHECK();
    const size_t smem = (size_t)nr * (32 + 12);
    if (smem > 48 * 1024) ICP_CUDA(cudaFuncSetAttribute(k_rbc_search_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t grid = div_up(m, EXACT_WARPS);
    const uint32_t cap = (uint32_t)ctx->sm_count * 8u;           // persistent-ish: a few waves, the staging of R is amortised
    if (grid > cap) grid = cap;
    k_rbc_search_exact<<<grid, EXACT_WARPS * 32, smem, ctx->stream>>>(d_Q, m, d_R, nr, d_Xp, d_O, d_N, psi, fg, fp, prune, d_NN_ID, d_NN,
                                                                      (unsigned long long *)d_evals);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
