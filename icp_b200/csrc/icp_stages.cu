// icp_stages.cu -- one CUDA kernel family per reference pipeline stage (sm_100a).
//
// Each launch_* function is the B200 replacement of one `<Stage>::run()` of the reference
// (/root/reference/src/ICP/algorithms.cpp); the arithmetic follows oracle/icp_oracle.cpp operation for
// operation (which in turn cites kernels/icp_kernels.cl).  Data layout in HBM: points are 8 x f32 AoS
// (two 16-byte vectors per point: one LDG.128 each), indices u32, per-stage partials SoA.
#include "icp_stages.cuh"
#include "icp_solve.cuh"

// =================================================================================================
// A1 / A2 sampling
// =================================================================================================
__global__ void k_get_lms(const float4 *__restrict__ in, float4 *__restrict__ out)
{
    // 2 threads per landmark, one 16-byte half each (icp_kernels.cl:62-76): row 49+3*gy, column 65+4*lx
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // [0, 128*256)
    if (t >= 128u * 256u) return;
    uint32_t gy = t >> 8, gx = t & 255u;
    uint32_t src = (49u + 3u * gy) * 1280u + (65u + 4u * (gx >> 1)) * 2u + (gx & 1u);
    out[t] = __ldg(in + src);
}

__global__ void k_get_reps(const float4 *__restrict__ lms, float4 *__restrict__ reps, uint32_t W, uint32_t nrx, uint32_t nry,
                           uint32_t sx, uint32_t sy)
{
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // 2 threads per representative
    if (t >= nrx * nry * 2u) return;
    uint32_t r = t >> 1, h = t & 1u;
    uint32_t gy = r / nrx, gx = r % nrx;
    uint32_t xi = gx * sx + (sx >> 1) - 1u;
    uint32_t yi = gy * sy + (sy >> 1) - 1u;
    reps[t] = __ldg(lms + ((size_t)yi * W + xi) * 2u + h);
}

// RGB-D -> pc8d (kinect_frame_grabber.cpp:246-263).  HBM-bound: 5 B in, 32 B out per pixel; grid-stride over the frame,
// one pixel per thread and trip, two 16-byte stores per pixel (a warp writes 1 KB contiguous).
__global__ void __launch_bounds__(256) k_rgbd_to_pc8d(const uint16_t *__restrict__ depth, const uint8_t *__restrict__ rgb,
                                                     uint32_t W, uint32_t n, float cx, float cy, float focal, float4 *__restrict__ out)
{
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x)
    {
        const uint32_t x = p % W, y = p / W;
        const float d = (float)__ldg(depth + p);
        const float r = (float)__ldg(rgb + 3u * p), g = (float)__ldg(rgb + 3u * p + 1u), b = (float)__ldg(rgb + 3u * p + 2u);
        float4 lo, hi;
        lo.x = __fdiv_rn(__fmul_rn(__fsub_rn((float)x, cx), d), focal);
        lo.y = __fdiv_rn(__fmul_rn(__fsub_rn((float)y, cy), d), focal);
        lo.z = d; lo.w = 1.f;
        hi.x = __fdiv_rn(r, 255.f); hi.y = __fdiv_rn(g, 255.f); hi.z = __fdiv_rn(b, 255.f); hi.w = 1.f;
        out[2u * p] = lo; out[2u * p + 1u] = hi;
    }
}

int launch_rgbd_to_pc8d(cudaStream_t st, int sm_count, const uint16_t *depth, const uint8_t *rgb, uint32_t W, uint32_t H, float focal, float *cloud)
{
    const uint32_t n = W * H;
    uint32_t blocks = div_up(n, 256);
    const uint32_t cap = (uint32_t)sm_count * 8u;            // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    k_rgbd_to_pc8d<<<blocks, 256, 0, st>>>(depth, rgb, W, n, (float)(W - 1) / 2.f, (float)(H - 1) / 2.f, focal, (float4 *)cloud);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

int launch_get_lms(cudaStream_t st, const float *cloud, float *lms)
{
    k_get_lms<<<128, 256, 0, st>>>((const float4 *)cloud, (float4 *)lms);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

int launch_get_reps(cudaStream_t st, const float *lms, uint32_t W, uint32_t H, uint32_t nr, float *reps)
{
    uint32_t nrx, nry;
    icp_rep_grid(nr, &nrx, &nry);
    if (nrx * nry != nr) ICP_CONFIG_FAIL("ICPReps", "The number of representatives has to be a power of 2");
    if (W % nrx || H % nry || W / nrx < 2 || H / nry < 2)
        ICP_CONFIG_FAIL("ICPReps", "The landmark grid is not divisible into the representative grid");
    k_get_reps<<<div_up(nr * 2, 128), 128, 0, st>>>((const float4 *)lms, (float4 *)reps, W, nrx, nry, W / nrx, H / nry);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// A4 / A15 transforms
// =================================================================================================
__global__ void k_transform_q(const float *__restrict__ M, const float *__restrict__ T, float *__restrict__ out, uint32_t m)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const float4 q = __ldg((const float4 *)T), t = __ldg((const float4 *)T + 1);
    pt8 p = ld_pt8(M, i);
    p.lo = transform_q_xyz(p.lo, q, t);
    st_pt8(out, i, p);
}

__global__ void k_transform_m(const float *__restrict__ M, const float *__restrict__ T, float *__restrict__ out, uint32_t m)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    pt8 p = ld_pt8(M, i);
    float v[4] = { p.lo.x, p.lo.y, p.lo.z, p.lo.w };
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const float4 row = __ldg((const float4 *)T + k);
        float tr[4] = { row.x, row.y, row.z, row.w };
        r[k] = solve::dot4_ip(tr, v);
    }
    p.lo.x = r[0]; p.lo.y = r[1]; p.lo.z = r[2];
    st_pt8(out, i, p);
}

int launch_transform_q(cudaStream_t st, const float *M, const float *T8, float *out, uint32_t m)
{
    k_transform_q<<<div_up(m, 256), 256, 0, st>>>(M, T8, out, m);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
int launch_transform_m(cudaStream_t st, const float *M, const float *T16, float *out, uint32_t m)
{
    k_transform_m<<<div_up(m, 256), 256, 0, st>>>(M, T16, out, m);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// RBC stage 1: nearest representative (A3 / A5).  S adjacent lanes share one point and scan nr/S
// representatives each from the shared-memory copy of R (broadcast LDS.128), then merge their
// (dist, id) with an ordered argmin (lower distance, then lower id) by warp shuffle.
// =================================================================================================
template <int S, bool TRANSFORM>
__global__ void __launch_bounds__(256) k_nearest_rep(const float *__restrict__ X, uint32_t n, const float *__restrict__ R, uint32_t nr,
                                                     float fg, float fp, const float *__restrict__ T, float *__restrict__ Xt_out,
                                                     uint32_t *__restrict__ rep_id)
{
    extern __shared__ float4 sR[];
    for (uint32_t i = threadIdx.x; i < nr * 2u; i += blockDim.x) sR[i] = __ldg((const float4 *)R + i);
    __syncthreads();

    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t qi = gt / S, c = gt % S;
    const bool valid = qi < n;
    pt8 q = ld_pt8(X, valid ? qi : 0u);
    if (TRANSFORM)
    {
        const float4 tq = __ldg((const float4 *)T), tt = __ldg((const float4 *)T + 1);
        q.lo = transform_q_xyz(q.lo, tq, tt);
        if (valid && c == 0 && Xt_out) st_pt8(Xt_out, qi, q);
    }
    const uint32_t chunk = (nr + S - 1) / S;
    const uint32_t r0 = c * chunk;
    const uint32_t r1 = min(nr, r0 + chunk);
    float best = CUDART_INF_F;
    uint32_t bi = r0;
#pragma unroll 4
    for (uint32_t r = r0; r < r1; ++r)
    {
        const float d = dist8(q.lo, q.hi, sR[2 * r], sR[2 * r + 1], fg, fp);
        if (d < best) { best = d; bi = r; }
    }
#pragma unroll
    for (int off = 1; off < S; off <<= 1)
    {
        const float od = __shfl_xor_sync(FULL_MASK, best, off);
        const uint32_t oi = __shfl_xor_sync(FULL_MASK, bi, off);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    if (valid && c == 0) rep_id[qi] = (best == CUDART_INF_F) ? 0u : bi;
}

template <int S>
static int launch_nearest_rep_s(cudaStream_t st, const float *X, uint32_t n, const float *R, uint32_t nr, float fg, float fp,
                                const float *T8, float *Xt_out, uint32_t *rep_id)
{
    const size_t smem = (size_t)nr * 32;
    const uint32_t grid = div_up((uint32_t)((uint64_t)n * S), 256);
    if (T8)
    {
        if (smem > 48 * 1024) ICP_CUDA(cudaFuncSetAttribute(k_nearest_rep<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nearest_rep<S, true><<<grid, 256, smem, st>>>(X, n, R, nr, fg, fp, T8, Xt_out, rep_id);
    }
    else
    {
        if (smem > 48 * 1024) ICP_CUDA(cudaFuncSetAttribute(k_nearest_rep<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_nearest_rep<S, false><<<grid, 256, smem, st>>>(X, n, R, nr, fg, fp, nullptr, nullptr, rep_id);
    }
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

int launch_nearest_rep(cudaStream_t st, int sm_count, const float *X, uint32_t n, const float *R, uint32_t nr,
                       float fg, float fp, const float *T8, float *Xt_out, uint32_t *rep_id)
{
    if (nr == 0 || nr > 7000) ICP_CONFIG_FAIL("RBC", "The number of representatives must be in [1, 7000]");
    // enough threads for >= 2 CTAs of 256 threads per SM
    const uint64_t want = (uint64_t)sm_count * 2 * 256;
    int S = 1;
    while (S < 8 && (uint64_t)n * S < want) S <<= 1;
    switch (S)
    {
        case 1: return launch_nearest_rep_s<1>(st, X, n, R, nr, fg, fp, T8, Xt_out, rep_id);
        case 2: return launch_nearest_rep_s<2>(st, X, n, R, nr, fg, fp, T8, Xt_out, rep_id);
        case 4: return launch_nearest_rep_s<4>(st, X, n, R, nr, fg, fp, T8, Xt_out, rep_id);
        default: return launch_nearest_rep_s<8>(st, X, n, R, nr, fg, fp, T8, Xt_out, rep_id);
    }
}

// =================================================================================================
// Stable counting sort by representative id (decision B4): histogram -> exclusive scan -> scatter.
// (The reference's RBC library does counts + scan + permute; scan semantics = scan_kernels.cl.)
// =================================================================================================
__global__ void __launch_bounds__(256) k_cs_count(const uint32_t *__restrict__ key, uint32_t n, uint32_t nr,
                                                  uint32_t *__restrict__ H, uint16_t *__restrict__ lrank)
{
    extern __shared__ uint16_t wcnt[];      // [8 warps][nr]
    const uint32_t tid = threadIdx.x, w = tid >> 5;
    for (uint32_t i = tid; i < 8u * nr; i += 256u) wcnt[i] = 0;
    __syncthreads();
    const uint32_t i = blockIdx.x * 256u + tid;
    const uint32_t k = (i < n) ? key[i] : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(FULL_MASK, k);
    const uint32_t lr = __popc(peers & lanemask_lt());
    if (i < n && lr == 0) wcnt[w * nr + k] = (uint16_t)__popc(peers);
    __syncthreads();
    if (i < n)
    {
        uint32_t base = 0;
        for (uint32_t w2 = 0; w2 < w; ++w2) base += wcnt[w2 * nr + k];
        lrank[i] = (uint16_t)(base + lr);
    }
    for (uint32_t r = tid; r < nr; r += 256u)
    {
        uint32_t s = 0;
#pragma unroll
        for (uint32_t w2 = 0; w2 < 8u; ++w2) s += wcnt[w2 * nr + r];
        H[(size_t)blockIdx.x * nr + r] = s;
    }
}

// one CTA: column-wise exclusive prefix of H over the CTAs, N = column totals, O = exclusive scan of N
__global__ void __launch_bounds__(1024) k_cs_scan(uint32_t *__restrict__ H, uint32_t nb, uint32_t nr,
                                                  uint32_t *__restrict__ N, uint32_t *__restrict__ O)
{
    __shared__ uint32_t warp_tot[32];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, w = tid >> 5;
    for (uint32_t r = tid; r < nr; r += 1024u)
    {
        uint32_t run = 0;
        for (uint32_t b = 0; b < nb; ++b)
        {
            const uint32_t v = H[(size_t)b * nr + r];
            H[(size_t)b * nr + r] = run;
            run += v;
        }
        N[r] = run;
    }
    __syncthreads();
    // exclusive scan of N: thread t owns the contiguous chunk [t*per, (t+1)*per)
    const uint32_t per = (nr + 1023u) / 1024u;
    const uint32_t b0 = tid * per;
    uint32_t sum = 0;
    for (uint32_t j = 0; j < per; ++j) if (b0 + j < nr) sum += N[b0 + j];
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t v = __shfl_up_sync(FULL_MASK, inc, d);
        if (lane >= (uint32_t)d) inc += v;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0)
    {
        uint32_t v = warp_tot[lane], iv = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t u = __shfl_up_sync(FULL_MASK, iv, d);
            if (lane >= (uint32_t)d) iv += u;
        }
        warp_tot[lane] = iv - v;   // exclusive
    }
    __syncthreads();
    uint32_t run = warp_tot[w] + inc - sum;
    for (uint32_t j = 0; j < per; ++j)
        if (b0 + j < nr) { O[b0 + j] = run; run += N[b0 + j]; }
}

__global__ void __launch_bounds__(256) k_cs_scatter(const uint32_t *__restrict__ key, uint32_t n, uint32_t nr,
                                                    const uint32_t *__restrict__ H, const uint16_t *__restrict__ lrank,
                                                    const uint32_t *__restrict__ O, uint32_t *__restrict__ perm,
                                                    const float *__restrict__ Xsrc, float *__restrict__ Xp)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = key[i];
    const uint32_t pos = O[k] + H[(size_t)blockIdx.x * nr + k] + lrank[i];
    if (perm) perm[pos] = i;
    if (Xp) st_pt8(Xp, pos, ld_pt8(Xsrc, i));
}

int launch_counting_sort(cudaStream_t st, const uint32_t *key, uint32_t n, uint32_t nr, SortScratch sc,
                         uint32_t *N, uint32_t *O, uint32_t *perm, const float *Xsrc, float *Xp)
{
    const uint32_t nb = div_up(n, 256);
    const size_t smem = (size_t)8 * nr * sizeof(uint16_t);
    if (smem > 48 * 1024) ICP_CUDA(cudaFuncSetAttribute(k_cs_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_cs_count<<<nb, 256, smem, st>>>(key, n, nr, sc.H, sc.lrank);
    ICP_LAUNCH_CHECK();
    k_cs_scan<<<1, 1024, 0, st>>>(sc.H, nb, nr, N, O);
    ICP_LAUNCH_CHECK();
    k_cs_scatter<<<nb, 256, 0, st>>>(key, n, nr, sc.H, sc.lrank, O, perm, Xsrc, Xp);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// RBC stage 2 (A5): thread p owns the p-th query in representative-sorted order and scans the list of
// its representative in X_p.  Consecutive threads share a representative, so the two LDG.128 per
// candidate are warp-uniform (one L1 transaction).  Strict '<' in ascending list position.
// =================================================================================================
__global__ void __launch_bounds__(128) k_rbc_stage2(const float *__restrict__ Q, const uint32_t *__restrict__ q_rep,
                                                    const uint32_t *__restrict__ qperm, uint32_t m, const float *__restrict__ Xp,
                                                    uint32_t n_db, const uint32_t *__restrict__ O, const uint32_t *__restrict__ N,
                                                    float fg, float fp, float *__restrict__ Qp, float *__restrict__ NN,
                                                    icp_dist_id *__restrict__ NNID, unsigned long long *__restrict__ evals)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cnt = 0;
    if (p < m)
    {
        const uint32_t i = qperm[p];
        const pt8 q = ld_pt8(Q, i);
        const uint32_t r = q_rep[i];
        const uint32_t o = O[r];
        cnt = N[r];
        float best = CUDART_INF_F;
        uint32_t bi = o;
        for (uint32_t k = o; k < o + cnt; ++k)
        {
            const pt8 x = ld_pt8(Xp, k);
            const float d = dist8(q.lo, q.hi, x.lo, x.hi, fg, fp);
            if (d < best) { best = d; bi = k; }
        }
        if (cnt == 0) bi = o ? o - 1u : 0u;        // empty list (cannot happen when R is a subset of X)
        if (bi >= n_db) bi = n_db - 1u;
        st_pt8(Qp, p, q);
        st_pt8(NN, p, ld_pt8(Xp, bi));
        icp_dist_id o2; o2.dist = best; o2.id = bi;
        NNID[p] = o2;
    }
    if (evals)
    {
        unsigned long long c = cnt;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_down_sync(FULL_MASK, c, d);
        if ((threadIdx.x & 31u) == 0 && c) atomicAdd(evals, c);
    }
}

int launch_rbc_stage2(cudaStream_t st, const float *Q, const uint32_t *q_rep, const uint32_t *qperm, uint32_t m,
                      const float *Xp, uint32_t n_db, const uint32_t *O, const uint32_t *N, float fg, float fp,
                      float *Qp, float *NN, icp_dist_id *NNID, unsigned long long *evals)
{
    k_rbc_stage2<<<div_up(m, 128), 128, 0, st>>>(Q, q_rep, qperm, m, Xp, n_db, O, N, fg, fp, Qp, NN, NNID, evals);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// Generic multi-level block reductions executed by ONE CTA (all threads participate).
//   QUAD = false: block = 128 consecutive values, slot = value              (icpGMean, reduce_sum_fd levels)
//   QUAD = true : block = 512 consecutive values, slot = (x+y)+z)+w of a float4   (reduce_sum_f)
// Levels repeat until one value remains; s0/s1 are ping-pong scratch buffers of >= ceil(cnt/per) values.
// =================================================================================================
template <typename Tv> __device__ __forceinline__ Tv warp_tree(Tv a, Tv b, Tv c, Tv d);
template <> __device__ __forceinline__ float warp_tree<float>(float a, float b, float c, float d) { return warp_tree128(a, b, c, d); }
template <> __device__ __forceinline__ double warp_tree<double>(double a, double b, double c, double d) { return warp_tree128_d(a, b, c, d); }

template <typename Tv> __device__ __forceinline__ Tv add_rn(Tv a, Tv b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }

template <typename Tv, bool QUAD>
__device__ Tv cta_reduce_levels(const Tv *src, uint32_t cnt, Tv *s0, Tv *s1)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const uint32_t per = QUAD ? 512u : 128u;
    const Tv *in = src;
    Tv *out = s0;
    while (true)
    {
        const uint32_t nb = (cnt + per - 1u) / per;
        for (uint32_t b = warp; b < nb; b += nwarps)
        {
            Tv e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t slot = lane + 32u * j;
                if (QUAD)
                {
                    const uint32_t idx = b * 512u + slot * 4u;
                    Tv v0 = idx < cnt ? in[idx] : Tv(0), v1 = idx + 1 < cnt ? in[idx + 1] : Tv(0);
                    Tv v2 = idx + 2 < cnt ? in[idx + 2] : Tv(0), v3 = idx + 3 < cnt ? in[idx + 3] : Tv(0);
                    e[j] = add_rn(add_rn(add_rn(v0, v1), v2), v3);
                }
                else
                {
                    const uint32_t idx = b * 128u + slot;
                    e[j] = idx < cnt ? in[idx] : Tv(0);
                }
            }
            const Tv s = warp_tree<Tv>(e[0], e[1], e[2], e[3]);
            if (lane == 0) out[b] = s;
        }
        __syncthreads();
        if (nb == 1) return out[0];
        cnt = nb;
        in = out;
        out = (out == s0) ? s1 : s0;
    }
}

// =================================================================================================
// A6 ICPWeights: w = 100/(100+dist); f32 tree per 128-element block (one warp each), then f64 levels.
// =================================================================================================
__global__ void __launch_bounds__(256) k_weights(const icp_dist_id *__restrict__ in, float *__restrict__ W,
                                                 float *__restrict__ bs, uint32_t n, uint32_t nb_padded)
{
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (gw >= nb_padded) return;
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const uint32_t idx = gw * 128u + lane + 32u * j;
        float w = 0.f;
        if (idx < n)
        {
            const float2 di = __ldg((const float2 *)in + idx);
            w = __fdiv_rn(100.f, __fadd_rn(100.f, di.x));
            W[idx] = w;
        }
        e[j] = w;
    }
    const float s = warp_tree128(e[0], e[1], e[2], e[3]);
    if (lane == 0) bs[gw] = s;
}

__global__ void __launch_bounds__(1024) k_weights_finish(const float *__restrict__ bs, uint32_t nb, uint32_t nb_padded,
                                                         double *s0, double *s1, double *__restrict__ sum_w)
{
    if (nb == 1) { if (threadIdx.x == 0) *sum_w = (double)bs[0]; return; }
    // reduce_sum_fd first step: one f64 per float4 of block sums
    const uint32_t nq = nb_padded / 4u;
    for (uint32_t k = threadIdx.x; k < nq; k += blockDim.x)
    {
        const float4 v = *((const float4 *)bs + k);
        s1[k] = __dadd_rn(__dadd_rn(__dadd_rn((double)v.x, (double)v.y), (double)v.z), (double)v.w);
    }
    __syncthreads();
    // s1 holds nq values (nq <= ceil(nb/4)+1 <= scratch size); levels ping-pong s0 <- s1 <- s0 ...
    const double r = cta_reduce_levels<double, false>(s1, nq, s0, s1 + nq);
    if (threadIdx.x == 0) *sum_w = r;
}

int launch_weights(cudaStream_t st, const icp_dist_id *in, float *W, double *sum_w, uint32_t n, float *scratchF, double *scratchD)
{
    const uint32_t nb = div_up(n, 128);
    const uint32_t nb_padded = (nb + 3u) & ~3u;
    k_weights<<<div_up(nb_padded * 32u, 256), 256, 0, st>>>(in, W, scratchF, n, nb_padded);
    ICP_LAUNCH_CHECK();
    k_weights_finish<<<1, 1024, 0, st>>>(scratchF, nb, nb_padded, scratchD, scratchD + reduce_scratch_elems(n), sum_w);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// A7 ICPMean<REGULAR|WEIGHTED>: per-element value = x/n  or  (float)(w/sum_w) * x, f32 tree per 128-point
// block (one warp per (set, block)), then levels over the block means (icpGMean).
// Block means are stored SoA: bm[(set*3 + c) * nb + b].
// =================================================================================================
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_mean_blocks(const float *__restrict__ F, const float *__restrict__ M,
                                                     const float *__restrict__ W, const double *__restrict__ sum_w,
                                                     float *__restrict__ bm, uint32_t n, uint32_t nb)
{
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (gw >= 2u * nb) return;
    const uint32_t set = gw / nb, b = gw % nb;
    const float *in = set ? M : F;
    const float fn = (float)n;
    double sw = 1.0;
    if (WEIGHTED) sw = *sum_w;
    float ex[4], ey[4], ez[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const uint32_t idx = b * 128u + lane + 32u * j;
        float x = 0.f, y = 0.f, z = 0.f;
        if (idx < n)
        {
            const float4 p = __ldg((const float4 *)in + (size_t)idx * 2);
            if (WEIGHTED)
            {
                const float wn = (float)__ddiv_rn((double)W[idx], sw);
                x = __fmul_rn(wn, p.x); y = __fmul_rn(wn, p.y); z = __fmul_rn(wn, p.z);
            }
            else { x = __fdiv_rn(p.x, fn); y = __fdiv_rn(p.y, fn); z = __fdiv_rn(p.z, fn); }
        }
        ex[j] = x; ey[j] = y; ez[j] = z;
    }
    const float sx = warp_tree128(ex[0], ex[1], ex[2], ex[3]);
    const float sy = warp_tree128(ey[0], ey[1], ey[2], ey[3]);
    const float sz = warp_tree128(ez[0], ez[1], ez[2], ez[3]);
    if (lane == 0)
    {
        bm[(set * 3u + 0u) * nb + b] = sx;
        bm[(set * 3u + 1u) * nb + b] = sy;
        bm[(set * 3u + 2u) * nb + b] = sz;
    }
}

// grid = 6 CTAs: one per (set, channel)
__global__ void __launch_bounds__(256) k_mean_finish(const float *__restrict__ bm, uint32_t nb, float *s0, float *s1, uint32_t sstride,
                                                     float *__restrict__ mean)
{
    const uint32_t row = blockIdx.x;               // set*3 + c
    float r;
    if (nb == 1) r = bm[row];
    else r = cta_reduce_levels<float, false>(bm + (size_t)row * nb, nb, s0 + (size_t)row * sstride, s1 + (size_t)row * sstride);
    if (threadIdx.x == 0)
    {
        mean[(row / 3u) * 4u + (row % 3u)] = r;
        if (row % 3u == 0) mean[(row / 3u) * 4u + 3u] = 0.f;
    }
}

int launch_mean(cudaStream_t st, const float *F, const float *M, const float *W, const double *sum_w, float *mean,
                uint32_t n, float *scratch)
{
    const uint32_t nb = div_up(n, 128);
    const uint32_t sstride = (uint32_t)reduce_scratch_elems(nb);
    float *bm = scratch;                                   // 6*nb
    float *s0 = bm + (size_t)6 * nb;                       // 6*sstride
    float *s1 = s0 + (size_t)6 * sstride;
    if (W) k_mean_blocks<true><<<div_up(2u * nb * 32u, 256), 256, 0, st>>>(F, M, W, sum_w, bm, n, nb);
    else k_mean_blocks<false><<<div_up(2u * nb * 32u, 256), 256, 0, st>>>(F, M, nullptr, nullptr, bm, n, nb);
    ICP_LAUNCH_CHECK();
    k_mean_finish<<<6, 256, 0, st>>>(bm, nb, s0, s1, sstride, mean);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// A8 ICPDevs
// =================================================================================================
__global__ void k_devs(const float *__restrict__ F, const float *__restrict__ M, const float *__restrict__ mean,
                       float4 *__restrict__ DF, float4 *__restrict__ DM, uint32_t n)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2u * n) return;
    const uint32_t set = t / n, i = t % n;
    const float4 p = __ldg((const float4 *)(set ? M : F) + (size_t)i * 2);
    const float4 mu = __ldg((const float4 *)mean + set);
    float4 d;
    d.x = __fsub_rn(p.x, mu.x); d.y = __fsub_rn(p.y, mu.y); d.z = __fsub_rn(p.z, mu.z); d.w = __fsub_rn(p.w, mu.w);
    (set ? DM : DF)[i] = d;
}

int launch_devs(cudaStream_t st, const float *F, const float *M, const float *mean, float *DF, float *DM, uint32_t n)
{
    k_devs<<<div_up(2u * n, 256), 256, 0, st>>>(F, M, mean, (float4 *)DF, (float4 *)DM, n);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// A9 ICPS<REGULAR|WEIGHTED>: G = ceil(m/4) threads, 4 strided pairs each, then Reduce<SUM> per row.
// =================================================================================================
template <bool WEIGHTED>
__global__ void __launch_bounds__(128) k_sij_partials(const float4 *__restrict__ DM, const float4 *__restrict__ DF,
                                                      const float *__restrict__ W, float *__restrict__ part, uint32_t m, uint32_t G, float c)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    float A[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) A[k] = 0.f;
    for (uint32_t pi = g; pi < m; pi += G)
    {
        const float4 dm = __ldg(DM + pi), df = __ldg(DF + pi);
        const float mp[3] = { __fmul_rn(c, dm.x), __fmul_rn(c, dm.y), __fmul_rn(c, dm.z) };
        const float fp[3] = { __fmul_rn(c, df.x), __fmul_rn(c, df.y), __fmul_rn(c, df.z) };
        const float ff = __fadd_rn(__fadd_rn(__fmul_rn(fp[0], fp[0]), __fmul_rn(fp[1], fp[1])), __fmul_rn(fp[2], fp[2]));
        const float mm = __fadd_rn(__fadd_rn(__fmul_rn(mp[0], mp[0]), __fmul_rn(mp[1], mp[1])), __fmul_rn(mp[2], mp[2]));
        if (WEIGHTED)
        {
            const float w = __ldg(W + pi);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(w, __fmul_rn(mp[a], fp[b])));
            A[9] = __fadd_rn(A[9], __fmul_rn(w, ff));
            A[10] = __fadd_rn(A[10], __fmul_rn(w, mm));
        }
        else
        {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    A[a * 3 + b] = __fadd_rn(A[a * 3 + b], __fmul_rn(mp[a], fp[b]));
            A[9] = __fadd_rn(A[9], ff);
            A[10] = __fadd_rn(A[10], mm);
        }
    }
#pragma unroll
    for (int k = 0; k < 11; ++k) part[(size_t)k * G + g] = A[k];
}

// Reduce<SUM,float>: one CTA per row
__global__ void __launch_bounds__(256) k_reduce_sum_rows(const float *__restrict__ in, uint32_t cols, float *s0, float *s1,
                                                         uint32_t sstride, float *__restrict__ out)
{
    const uint32_t row = blockIdx.x;
    const float r = cta_reduce_levels<float, true>(in + (size_t)row * cols, cols, s0 + (size_t)row * sstride, s1 + (size_t)row * sstride);
    if (threadIdx.x == 0) out[row] = r;
}

int launch_reduce_sum_f(cudaStream_t st, const float *in, uint32_t cols, uint32_t rows, float *out, float *scratch)
{
    const uint32_t sstride = div_up(cols, 512) + 8;
    k_reduce_sum_rows<<<rows, 256, 0, st>>>(in, cols, scratch, scratch + (size_t)rows * sstride, sstride, out);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

int launch_sij(cudaStream_t st, const float *DM, const float *DF, const float *W, float *S11, uint32_t m, float c,
               float *partials, float *scratch)
{
    const uint32_t G = div_up(m, 4);
    if (W) k_sij_partials<true><<<div_up(G, 128), 128, 0, st>>>((const float4 *)DM, (const float4 *)DF, W, partials, m, G, c);
    else k_sij_partials<false><<<div_up(G, 128), 128, 0, st>>>((const float4 *)DM, (const float4 *)DF, nullptr, partials, m, G, c);
    ICP_LAUNCH_CHECK();
    return launch_reduce_sum_f(st, partials, G, 11, S11, scratch);
}

// =================================================================================================
// A10 / A11 / A12 single-thread solvers
// =================================================================================================
__global__ void k_power_method(const float *__restrict__ S, const float *__restrict__ mean, float *__restrict__ Tk)
{
    if (threadIdx.x != 0) return;
    float s[11], mu[8], tk[8];
    for (int i = 0; i < 11; ++i) s[i] = S[i];
    for (int i = 0; i < 8; ++i) mu[i] = mean[i];
    solve::power_method(s, mu, tk);
    for (int i = 0; i < 8; ++i) Tk[i] = tk[i];
}

__global__ void k_svd_solve(const float *__restrict__ S, const float *__restrict__ mean, float *__restrict__ Tk, float *__restrict__ Rk)
{
    if (threadIdx.x != 0) return;
    float s[11], mu[8], tk[8], rk[9];
    for (int i = 0; i < 11; ++i) s[i] = S[i];
    for (int i = 0; i < 8; ++i) mu[i] = mean[i];
    solve::svd_solve(s, mu, tk, rk);
    for (int i = 0; i < 8; ++i) Tk[i] = tk[i];
    if (Rk) for (int i = 0; i < 9; ++i) Rk[i] = rk[i];
}

__global__ void k_accumulate(DevState *state, const float *__restrict__ Tk, const float *__restrict__ Rk, float *__restrict__ T,
                             int check, uint32_t max_iterations, double angle_thr, double trans_thr)
{
    if (threadIdx.x != 0) return;
    float tk[8], rk[9], t8[8];
    for (int i = 0; i < 8; ++i) tk[i] = Tk[i];
    if (Rk) for (int i = 0; i < 9; ++i) rk[i] = Rk[i];
    solve::accumulate(state, tk, Rk ? rk : nullptr, t8);
    for (int i = 0; i < 8; ++i) T[i] = t8[i];
    if (check) solve::check_convergence(state, max_iterations, angle_thr, trans_thr);
    else state->k = state->k + 1;
}

int launch_power_method(cudaStream_t st, const float *S11, const float *mean, float *Tk8)
{
    k_power_method<<<1, 32, 0, st>>>(S11, mean, Tk8);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
int launch_svd_solve(cudaStream_t st, const float *S11, const float *mean, float *Tk8, float *Rk9)
{
    k_svd_solve<<<1, 32, 0, st>>>(S11, mean, Tk8, Rk9);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
int launch_accumulate(cudaStream_t st, DevState *state, const float *Tk8, const float *Rk9, float *T8,
                      int check, uint32_t max_iterations, double angle_thr, double trans_thr)
{
    k_accumulate<<<1, 32, 0, st>>>(state, Tk8, Rk9, T8, check, max_iterations, angle_thr, trans_thr);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// =================================================================================================
// Reduce<MIN,float>, Reduce<MAX,uint>, Scan<int>  (utility classes of the reference; min/max are order independent)
// =================================================================================================
template <typename Tv, bool IS_MIN>
__global__ void __launch_bounds__(256) k_reduce_minmax_rows(const Tv *__restrict__ in, uint32_t cols, Tv *__restrict__ out)
{
    __shared__ Tv sm[8];
    const Tv *row = in + (size_t)blockIdx.x * cols;
    Tv v = row[0];
    for (uint32_t c = threadIdx.x; c < cols; c += blockDim.x)
    {
        const Tv x = row[c];
        if (IS_MIN) v = (x < v) ? x : v; else v = (x > v) ? x : v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const Tv o = __shfl_down_sync(FULL_MASK, v, d);
        if (IS_MIN) v = (o < v) ? o : v; else v = (o > v) ? o : v;
    }
    if ((threadIdx.x & 31u) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < 8; ++w) { const Tv o = sm[w]; if (IS_MIN) v = (o < v) ? o : v; else v = (o > v) ? o : v; }
        out[blockIdx.x] = v;
    }
}

int launch_reduce_min_f(cudaStream_t st, const float *in, uint32_t cols, uint32_t rows, float *out)
{
    k_reduce_minmax_rows<float, true><<<rows, 256, 0, st>>>(in, cols, out);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
int launch_reduce_max_ui(cudaStream_t st, const uint32_t *in, uint32_t cols, uint32_t rows, uint32_t *out)
{
    k_reduce_minmax_rows<uint32_t, false><<<rows, 256, 0, st>>>(in, cols, out);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}

// one CTA per row, 256 threads, chunked carry (integer add: association free)
__global__ void __launch_bounds__(256) k_scan_rows(const int32_t *__restrict__ in, uint32_t cols, int inclusive, int32_t *__restrict__ out)
{
    __shared__ int32_t warp_tot[8];
    __shared__ int32_t carry_s;
    const int32_t *row = in + (size_t)blockIdx.x * cols;
    int32_t *orow = out + (size_t)blockIdx.x * cols;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < cols; base += 256u)
    {
        const uint32_t c = base + threadIdx.x;
        const int32_t v = c < cols ? row[c] : 0;
        int32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            const int32_t u = __shfl_up_sync(FULL_MASK, inc, d);
            if (lane >= (uint32_t)d) inc += u;
        }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        int32_t wbase = 0;
        for (uint32_t w2 = 0; w2 < w; ++w2) wbase += warp_tot[w2];
        const int32_t carry = carry_s;
        const int32_t res = carry + wbase + inc;
        if (c < cols) orow[c] = inclusive ? res : res - v;
        __syncthreads();
        if (threadIdx.x == 255) carry_s = res;
        __syncthreads();
    }
}

int launch_scan_i(cudaStream_t st, const int32_t *in, uint32_t cols, uint32_t rows, int inclusive, int32_t *out)
{
    k_scan_rows<<<rows, 256, 0, st>>>(in, cols, inclusive, out);
    ICP_LAUNCH_CHECK();
    return ICP_OK;
}
