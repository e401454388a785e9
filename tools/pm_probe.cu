// Micro-benchmark + equivalence check of the power-method flavours (single thread / warp-cooperative / redundant lanes).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
//        -I icp_b200/csrc -I include tools/pm_probe.cu -o build/pm_probe
#define PM_PROBE 1
#include "icp_solve.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

__global__ void k_probe(const float *S, const float *mu, int n_cases, float *out, long long *cyc, int *iters)
{
    __shared__ __align__(16) float ring[32][4];
    const int lane = threadIdx.x;
    for (int c = 0; c < n_cases; ++c)
    {
        float s11[11], m8[8], tk[8];
        for (int i = 0; i < 11; ++i) s11[i] = S[c * 11 + i];
        for (int i = 0; i < 8; ++i) m8[i] = mu[c * 8 + i];
        for (int v = 0; v < 3; ++v)
        {
            __syncwarp();
            long long t0 = clock64();
            int it = 0;
            if (v == 0) { if (lane == 0) it = solve::power_method(s11, m8, tk); }
            else if (v == 1) it = solve::power_method_warp(s11, m8, tk, ring);
            else it = solve::power_method_warp2(s11, m8, tk, ring);
            __syncwarp();
            long long t1 = clock64();
            if (lane == 0)
            {
                for (int i = 0; i < 8; ++i) out[(c * 3 + v) * 8 + i] = tk[i];
                cyc[c * 3 + v] = t1 - t0; iters[c * 3 + v] = it;
            }
        }
    }
}

// exhaustive-style check of the branch-free fast paths against the library operations: n random operands per thread
__global__ void k_check_fast(unsigned long long seed, int n, unsigned long long *bad)
{
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1);
    unsigned long long nb_s = 0, nb_d = 0, used_s = 0, used_d = 0;
    for (int i = 0; i < n; ++i)
    {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const uint32_t ua = (uint32_t)x, ub = (uint32_t)(x >> 32);
        // sqrt: any positive float in the fast-path range
        float v = __uint_as_float(ua & 0x7FFFFFFFu);
        if ((__float_as_uint(v) - 0x0d000000u) <= 0x727fffffu)
        {
            ++used_s;
            if (__float_as_uint(solve::sqrt_fast_path(v)) != __float_as_uint(__fsqrt_rn(v))) ++nb_s;
        }
        // div: both operands forced into the band (exponent 64..190), random signs and mantissas
        const uint32_t ea = 64u + ((ua >> 23) & 0xFFu) % 127u, eb = 64u + ((ub >> 23) & 0xFFu) % 127u;
        const float a = __uint_as_float((ua & 0x807FFFFFu) | (ea << 23)), b = __uint_as_float((ub & 0x007FFFFFu) | (eb << 23));
        if (solve::div_band(a) && solve::div_band(b))
        {
            ++used_d;
            if (__float_as_uint(solve::div_fast_path(a, b, solve::rcp_refined(b))) != __float_as_uint(__fdiv_rn(a, b))) ++nb_d;
        }
    }
    atomicAdd(bad, nb_s); atomicAdd(bad + 1, nb_d); atomicAdd(bad + 2, used_s); atomicAdd(bad + 3, used_d);
}

// dependent-issue latencies (one warp, chains of 512 dependent operations)
template <int OP> __global__ void k_lat(float a, float b, float *out, long long *cyc)
{
    float x = a + threadIdx.x * 1e-7f;
    float xs1 = b, xs2 = a, xs3 = b; const float x0s = 0.25f * b;
    __syncwarp();
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 512; ++i)
    {
        if (OP == 0) x = __fadd_rn(x, b);
        else if (OP == 1) x = __fmul_rn(x, b);
        else if (OP == 2) x = __fmaf_rn(x, b, b);
        else if (OP == 3) x = solve::mufu_rcp(x);
        else if (OP == 4) x = solve::mufu_rsq(x);
        else if (OP == 5) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 3, 4);
        else if (OP == 6) x = __fdiv_rn(b, x);
        else if (OP == 7) x = __fsqrt_rn(x);
        else if (OP == 8) x = solve::div_fast_path(b, x, solve::rcp_refined(x));
        else if (OP == 9) x = solve::sqrt_fast_path(x);
        else if (OP == 10) { float v[4] = { x, b, a, b }; solve::pm_normalize_fast(v); x = __fadd_rn(v[0], v[3]); }
        else if (OP == 11) { float v[4] = { x, b, a, b }; solve::pm_normalize(v); x = __fadd_rn(v[0], v[3]); }
        else if (OP == 12) { double d = (double)x; d = __dadd_rn(d, (double)b); x = (float)d; }
        else if (OP == 13) { double d = (double)x; d = __dadd_rn(d, (double)b); d = __dadd_rn(d, (double)b); d = __dadd_rn(d, (double)b); d = __dadd_rn(d, (double)b); x = (float)d; }
        else if (OP == 15) { const float N[16] = { a, b, b, x0s, b, a, x0s, b, b, x0s, a, b, x0s, b, b, a }; float v[4] = { x, xs1, xs2, xs3 }; uint32_t bad = 0u;
                             solve::pm_trip_fast(N, v, bad); x = bad ? 0.f : v[0]; xs1 = v[1]; xs2 = v[2]; xs3 = v[3]; }
        else if (OP == 14) { float y[4] = { x, b, a, b }, z[4] = { a, x, b, a }; x = __fadd_rn(solve::pm_distance(y, z), 1.f); }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { out[OP] = x + xs1 + xs2 + xs3; cyc[OP] = t1 - t0; }
}

int main()
{
    {
        float *dO; long long *dC, hC[16];
        cudaMalloc(&dO, 64); cudaMalloc(&dC, 128);
        k_lat<0><<<1, 32>>>(1.0f, 1.0001f, dO, dC); k_lat<1><<<1, 32>>>(1.0f, 1.0001f, dO, dC); k_lat<2><<<1, 32>>>(1.0f, 0.5f, dO, dC);
        k_lat<3><<<1, 32>>>(1.3f, 1.0001f, dO, dC); k_lat<4><<<1, 32>>>(1.3f, 1.0001f, dO, dC); k_lat<5><<<1, 32>>>(1.0f, 1.0001f, dO, dC);
        k_lat<6><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<7><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<8><<<1, 32>>>(1.3f, 1.7f, dO, dC);
        k_lat<9><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<10><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<11><<<1, 32>>>(1.3f, 1.7f, dO, dC);
        k_lat<12><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<13><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<14><<<1, 32>>>(1.3f, 1.7f, dO, dC); k_lat<15><<<1, 32>>>(1.3f, 1.7f, dO, dC);
        cudaMemcpy(hC, dC, 128, cudaMemcpyDeviceToHost);
        const char *nm[16] = { "FADD", "FMUL", "FFMA", "MUFU.RCP", "MUFU.RSQ", "SHFL", "__fdiv_rn", "__fsqrt_rn", "div fast path (+rcp refine)", "sqrt fast path",
                               "pm_normalize_fast + 1 FADD", "pm_normalize + 1 FADD", "F2F.64 + DADD + F2F.32", "F2F.64 + 4 DADD + F2F.32", "pm_distance + 1 FADD", "pm_trip_fast" };
        for (int i = 0; i < 16; ++i) printf("latency %-28s %6.1f cycles\n", nm[i], (double)hC[i] / 512.0);
    }
    {
        unsigned long long *dbad, hbad[4];
        cudaMalloc(&dbad, 32); cudaMemset(dbad, 0, 32);
        k_check_fast<<<148 * 8, 256>>>(12345ull, 8192, dbad);
        cudaMemcpy(hbad, dbad, 32, cudaMemcpyDeviceToHost);
        printf("fast-path check: sqrt %llu mismatches of %llu, div %llu mismatches of %llu\n", hbad[0], hbad[2], hbad[1], hbad[3]);
        if (hbad[0] || hbad[1]) { printf("MISMATCH in fast paths\n"); return 2; }
    }
    const int NC = 96;
    float hS[NC * 11], hM[NC * 8];
    const float kS[11] = {0.00168053f, 0.000131408f, -0.000775179f, 0.000156595f, 0.00102674f, -0.000563479f,
                          -0.000722137f, -0.000559463f, 0.00246661f, 0.00521271f, 0.00515292f};
    const float kM[8] = {-33.9694f, -17.6421f, 1494.22f, 0.f, -44.8322f, -19.3835f, 1485.93f, 0.f};
    srand(7);
    for (int c = 0; c < NC; ++c)
    {
        for (int i = 0; i < 11; ++i)
        {
            float r = (float)rand() / RAND_MAX - 0.5f;
            hS[c * 11 + i] = c == 0 ? kS[i] : (c % 3 == 1 ? kS[i] * (1.f + 0.3f * r) : (c % 3 == 2 ? (i < 9 ? r * 1e-3f : 5e-3f + 1e-4f * r) : -kS[i] + 1e-4f * r));
        }
        if (c % 3 == 0 && c) { hS[c * 11 + 9] = fabsf(hS[c * 11 + 9]) + 1e-3f; hS[c * 11 + 10] = fabsf(hS[c * 11 + 10]) + 1e-3f; }
        if (c == 4 || c == 40)           // symmetric, traceless S: row 3 of N is exactly zero => zero component => library fallback path
        {
            float *S = hS + c * 11;
            S[3] = S[1]; S[6] = S[2]; S[7] = S[5]; S[8] = -(S[0] + S[4]);
        }
        if (c == 5) for (int i = 0; i < 9; ++i) hS[c * 11 + i] *= 1e-20f;      // tiny: sum of squares underflows the band
        for (int i = 0; i < 8; ++i) hM[c * 8 + i] = kM[i] * (1.f + 0.01f * c);
    }
    float *dS, *dM, *dO; long long *dC; int *dI;
    cudaMalloc(&dS, sizeof(hS)); cudaMalloc(&dM, sizeof(hM)); cudaMalloc(&dO, NC * 3 * 8 * 4); cudaMalloc(&dC, NC * 3 * 8); cudaMalloc(&dI, NC * 3 * 4);
    cudaMemcpy(dS, hS, sizeof(hS), cudaMemcpyHostToDevice); cudaMemcpy(dM, hM, sizeof(hM), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) k_probe<<<1, 32>>>(dS, dM, NC, dO, dC, dI);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    float hO[NC * 3 * 8]; long long hC[NC * 3]; int hI[NC * 3];
    cudaMemcpy(hO, dO, sizeof(hO), cudaMemcpyDeviceToHost); cudaMemcpy(hC, dC, sizeof(hC), cudaMemcpyDeviceToHost); cudaMemcpy(hI, dI, sizeof(hI), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int c = 0; c < NC; ++c)
    {
        const bool eq1 = !memcmp(hO + (c * 3) * 8, hO + (c * 3 + 1) * 8, 32) && hI[c * 3] == hI[c * 3 + 1];
        const bool eq2 = !memcmp(hO + (c * 3) * 8, hO + (c * 3 + 2) * 8, 32) && hI[c * 3] == hI[c * 3 + 2];
        bad += !eq1 + !eq2;
        if (c < 12 || !eq1 || !eq2) printf("case %2d: trips %4d | cycles single %7lld (%5.0f/trip) warp %7lld (%5.0f/trip) warp2 %7lld (%5.0f/trip) | equal %d %d\n", c, hI[c * 3],
               hC[c * 3], (double)hC[c * 3] / hI[c * 3], hC[c * 3 + 1], (double)hC[c * 3 + 1] / hI[c * 3], hC[c * 3 + 2], (double)hC[c * 3 + 2] / hI[c * 3], eq1, eq2);
    }
    {
        long long dbg[64];
        cudaMemcpyFromSymbol(dbg, solve::g_pm_dbg, sizeof(dbg));
        printf("last case, warp2 stamps (cycles): ");
        for (int i = 0; i + 2 < 30; i += 3) printf("[trips %lld test %lld gap %lld] ", dbg[i + 1] - dbg[i], dbg[i + 2] - dbg[i + 1], dbg[i + 3] - dbg[i + 2]);
        printf("\n");
    }
    printf("%s\n", bad ? "MISMATCH" : "all flavours bit-identical");
    return bad != 0;
}
