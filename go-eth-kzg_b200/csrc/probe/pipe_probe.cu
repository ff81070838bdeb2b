// pipe_probe -- standalone microbenchmarks behind the roofline denominators and the FP64-pipe design decision.
// Not part of libkzgb200.so.  Build + run: scripts/run_pipe_probe.sh (nvcc -gencode arch=compute_100a,code=sm_100a).
// Prints one JSON object.  Every throughput figure is device-wide lane-operations per second, CUDA-event timed,
// best of 5 after a warm-up launch.
#include "../fp64.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace kzg;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- raw instruction probes: NCH independent chains per thread -------------------------------------------
// MODE 0 mad.lo.u32, 1 mad.wide.u32 (no carry), 2 mad.lo.cc/madc.hi.cc pairs (the fused IMAD.WIDE.X form the field code uses),
// 3 fma.rn.f64, 4 wide + dfma interleaved 1:1, 5 wide + dfma 1:2, 6 wide + iadd3 1:2, 7 wide + dfma warp-specialised (even/odd warps)
template <int MODE, int NCH> __global__ void __launch_bounds__(256) k_raw(uint32_t *out, int iters, uint32_t seed) {
    uint32_t m = blockIdx.x * 2654435761u + 12345u + seed;
    unsigned long long a[NCH];
    double d[NCH];
    uint32_t x[NCH], y[NCH];
    const double dm = 1.0 + 1e-9 * (seed & 7);
#pragma unroll
    for (int k = 0; k < NCH; ++k) { a[k] = (unsigned long long)seed * (2 * k + 3) + threadIdx.x; d[k] = 1.0 + k + threadIdx.x * 1e-3; x[k] = seed * (2 * k + 5) + threadIdx.x; y[k] = x[k] ^ 5u; }
    const bool wide_warp = ((threadIdx.x >> 5) & 1) == 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) {
                if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m), "r"(x[(k + 1) % NCH]));
                if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 6 || (MODE == 7 && wide_warp)) {
                    // mad.lo.cc + madc.hi on one limb pair: ptxas fuses the pair into ONE IMAD.WIDE.U32 with a 64-bit addend (checked in SASS);
                    // a bare mad.wide.u32 with a 64-bit accumulator is split into IMAD.WIDE + IADD3 + IADD3.X instead
                    uint32_t lo = (uint32_t)a[k], hi = (uint32_t)(a[k] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"((uint32_t)(a[(k + 1) % NCH] >> 32)), "r"(m));
                    a[k] = ((unsigned long long)hi << 32) | lo;
                }
                if (MODE == 3 || MODE == 4 || MODE == 5) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[k]) : "d"(dm));
                if (MODE == 5) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[k]) : "d"(dm));
                if (MODE == 6) { asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(m)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(y[k]) : "r"(m)); }
                if (MODE == 7 && !wide_warp) asm volatile("fma.rn.f64 %0, %0, %1, %0;" : "+d"(d[k]) : "d"(dm));
            }
        }
    }
    unsigned long long r = 0;
#pragma unroll
    for (int k = 0; k < NCH; ++k) r ^= a[k] ^ (unsigned long long)__double_as_longlong(d[k]) ^ x[k] ^ y[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(r ^ (r >> 32));
}

// ---- field-multiplication probes -----------------------------------------------------------------------
// MODE 0: Fp::mul (IMAD) chain, 1: fp_mul_f64 chain, 2: both in one thread (two independent chains), 3: warp-specialised,
// 4: Fp::sqr chain, 5: fp_sqr_f64 chain, 6: two IMAD chains per thread (ILP reference for mode 2)
template <int MODE> __global__ void __launch_bounds__(128) k_fpmul(const uint32_t *in, uint32_t *out, int iters) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    Fp x, y, z;
#pragma unroll
    for (int k = 0; k < 12; ++k) { x.v[k] = in[gid * 24 + k]; y.v[k] = in[gid * 24 + 12 + k]; }
    z = y;
    const bool imad_warp = ((threadIdx.x >> 5) & 1) == 0;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) x = Fp::mul(x, y);
        if (MODE == 1) x = fp_mul_f64(x, y);
        if (MODE == 2) { x = Fp::mul(x, y); z = fp_mul_f64(z, y); }
        if (MODE == 3) { if (imad_warp) x = Fp::mul(x, y); else x = fp_mul_f64(x, y); }
        if (MODE == 4) x = Fp::sqr(x);
        if (MODE == 5) x = fp_sqr_f64(x);
        if (MODE == 6) { x = Fp::mul(x, y); z = Fp::mul(z, y); }
    }
#pragma unroll
    for (int k = 0; k < 12; ++k) out[gid * 12 + k] = x.v[k] ^ ((MODE == 2 || MODE == 6) ? z.v[k] : 0u);
}

// correctness: out = (a*b by IMAD) xor (a*b by FP64) ... must be all zero; same for squares
__global__ void k_check(const uint32_t *in, uint32_t *bad, int n) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    Fp x, y;
#pragma unroll
    for (int k = 0; k < 12; ++k) { x.v[k] = in[gid * 24 + k]; y.v[k] = in[gid * 24 + 12 + k]; }
    // bring arbitrary 384-bit inputs below p: two conditional subtractions are not enough in general, so mask the top limb
    x.v[11] &= 0x0fffffffu; y.v[11] &= 0x0fffffffu;
    Fp::final_sub(x.v); Fp::final_sub(y.v);
    if (gid == 0) { for (int k = 0; k < 12; ++k) { x.v[k] = FP_MOD[k]; y.v[k] = FP_MOD[k]; } x.v[0] -= 1; y.v[0] -= 1; }   // (p-1)^2
    if (gid == 1) { x = Fp::zero(); }
    for (int r = 0; r < 8; ++r) {
        Fp a = Fp::mul(x, y), b = fp_mul_f64(x, y);
        Fp c = Fp::sqr(x), d = fp_sqr_f64(x);
        if (!Fp::eq(a, b)) atomicAdd(&bad[0], 1u);
        if (!Fp::eq(c, d)) atomicAdd(&bad[1], 1u);
        x = a; y = c;
    }
}

template <class F> static double best_ms(F launch) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0)); launch(rep); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, clk);
    uint32_t *d_out; CK(cudaMalloc(&d_out, (size_t)sms * 16 * 256 * 48));
    // raw probes
    printf(" \"raw_T_lane_ops_per_s\": {\n");
    struct { const char *name; int mode; double ops_per_inner; } raws[] = {
        {"mad_lo", 0, 1}, {"imad_wide (mad.lo.cc+madc.hi pair)", 2, 1}, {"dfma", 3, 1},
        {"wide+dfma 1:1 (each)", 4, 1}, {"wide+2dfma (wide count)", 5, 1}, {"wide+2alu (wide count)", 6, 1}, {"wide|dfma warp-specialised (each warp class)", 7, 1}};
    for (int wpb : {4, 8}) for (int bps : {1, 2, 4}) {
        const int threads = wpb * 32, blocks = sms * bps, iters = 2048;
        for (auto &r : raws) {
            double ms = best_ms([&](int rep) {
                switch (r.mode) {
                    case 0: k_raw<0, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 1: k_raw<1, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 2: k_raw<2, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 3: k_raw<3, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 4: k_raw<4, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 5: k_raw<5, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    case 6: k_raw<6, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                    default: k_raw<7, 8><<<blocks, threads>>>(d_out, iters, 17u + rep); break;
                }
            });
            double ops = (double)blocks * threads * iters * 8.0 * 8.0;
            if (r.mode == 7) ops *= 0.5;
            printf("  \"%s @%dw/blk x%d blk/SM\": %.3f,\n", r.name, wpb, bps, ops / (ms * 1e-3) / 1e12);
        }
    }
    printf("  \"_\": 0},\n");
    // field multiplication probes
    const int n = sms * 8 * 128;
    std::vector<uint32_t> h((size_t)n * 24);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (auto &w : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; w = (uint32_t)(s >> 16); }
    for (int i = 0; i < n; ++i) { h[(size_t)i * 24 + 11] &= 0x0fffffffu; h[(size_t)i * 24 + 23] &= 0x0fffffffu; }   // < 2^380 < p
    uint32_t *d_in, *d_bad; CK(cudaMalloc(&d_in, h.size() * 4)); CK(cudaMalloc(&d_bad, 8));
    CK(cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_bad, 0, 8));
    k_check<<<(n + 127) / 128, 128>>>(d_in, d_bad, n);
    uint32_t bad[2]; CK(cudaMemcpy(bad, d_bad, 8, cudaMemcpyDeviceToHost));
    printf(" \"fp64_vs_imad_mismatches\": {\"mul\": %u, \"sqr\": %u, \"cases\": %d},\n", bad[0], bad[1], n * 8);
    printf(" \"fp_mul_G_per_s\": {\n");
    const char *names[] = {"imad mul", "f64 mul", "imad+f64 in one thread (total muls)", "imad|f64 warp-specialised (total)", "imad sqr", "f64 sqr", "2x imad in one thread (total)"};
    for (int bps : {2, 4, 8}) {
        const int blocks = sms * bps, iters = 256;
        for (int mode = 0; mode < 7; ++mode) {
            double ms = best_ms([&](int) {
                switch (mode) {
                    case 0: k_fpmul<0><<<blocks, 128>>>(d_in, d_out, iters); break;
                    case 1: k_fpmul<1><<<blocks, 128>>>(d_in, d_out, iters); break;
                    case 2: k_fpmul<2><<<blocks, 128>>>(d_in, d_out, iters); break;
                    case 3: k_fpmul<3><<<blocks, 128>>>(d_in, d_out, iters); break;
                    case 4: k_fpmul<4><<<blocks, 128>>>(d_in, d_out, iters); break;
                    case 5: k_fpmul<5><<<blocks, 128>>>(d_in, d_out, iters); break;
                    default: k_fpmul<6><<<blocks, 128>>>(d_in, d_out, iters); break;
                }
            });
            double muls = (double)blocks * 128 * iters * ((mode == 2 || mode == 6) ? 2.0 : 1.0);
            printf("  \"%s @%d blk/SM\": %.3f,\n", names[mode], bps, muls / (ms * 1e-3) / 1e9);
        }
    }
    printf("  \"_\": 0}}\n");
    return 0;
}
