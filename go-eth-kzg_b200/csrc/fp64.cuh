// Fp Montgomery multiplication on the FP64 pipe (DFMA), bit-exact with Fp::mul of mont.cuh.
//
// Why: IMAD.WIDE issues only on the fmaheavy sub-pipe (32 lanes/clk/SM), and every kernel of the proving
// paths sits on that one pipe while the B200's full-rate FP64 pipe (64 DFMA lanes/clk/SM) idles.  A double
// holds integers < 2^53 exactly, so with 16 limbs of 24 bits (16 * 24 = 384 = the Montgomery radix of the
// 12 x 32-bit form: R is the SAME, operands need no domain change) a column of the schoolbook product
// -- at most 16 a_i*b_j + 16 q_i*p_j products of < 2^48 plus a carry < 2^29 -- stays below 2^53 and every
// fma is exact: one DFMA per limb product, no hi/lo splitting, no integer accumulation.
//
// Row i (operand scanning, reduction interleaved, 24-bit digits):
//     t[j] += a[j] * b[i]                      16 DFMA
//     q     = ((t[0] mod 2^24) * (-p^-1)) mod 2^24      (mod 2^24 by the 2^76 round-toward-zero trick)
//     t[j] += q * p[j]                         16 DFMA      -> t[0] = 0 mod 2^24
//     t[1] += t[0] / 2^24 ; shift down one limb
// After 16 rows t < 2p; limbs are carry-normalised, repacked to 12 x 32 bits and conditionally reduced.
// (Python model with the bound check: tests/test_fp64_model.py.)
#pragma once
#include "field.cuh"

namespace kzg {

// p in 24-bit limbs, as doubles
__device__ __constant__ const double FP64_P[16] = {
    16755371.0 /*0xffaaab*/, 16777215.0 /*0xffffff*/, 16759294.0 /*0xffb9fe*/, 11621375.0 /*0xb153ff*/,
    11272190.0 /*0xabfffe*/, 16131102.0 /*0xf6241e*/, 10548912.0 /*0xa0f6b0*/, 6762706.0 /*0x6730d2*/,
    8721087.0 /*0x8512bf*/, 4949235.0 /*0x4b84f3*/, 14115959.0 /*0xd76477*/, 4410284.0 /*0x434bac*/,
    1812406.0 /*0x1ba7b6*/, 15112779.0 /*0xe69a4b*/, 15350143.0 /*0xea397f*/, 1704209.0 /*0x1a0111*/};
#define KZG_FP64_PINV 16580605.0        /* -p^-1 mod 2^24 = 0xfcfffd */
#define KZG_FP64_M76 75557863725914323419136.0   /* 2^76: ulp 2^24 */
#define KZG_FP64_2P52 4503599627370496.0
#define KZG_FP64_INV24 5.9604644775390625e-08   /* 2^-24 */

// 12 x u32 -> 16 x 24-bit limbs as exact doubles
__device__ __forceinline__ void fp_to_f64(double *d, const uint32_t *v) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t w0 = v[3 * g], w1 = v[3 * g + 1], w2 = v[3 * g + 2];
        uint32_t l0 = w0 & 0xffffffu;
        uint32_t l1 = __funnelshift_r(w0, w1, 24) & 0xffffffu;
        uint32_t l2 = __funnelshift_r(w1, w2, 16) & 0xffffffu;
        uint32_t l3 = w2 >> 8;
        d[4 * g + 0] = __hiloint2double(0x43300000, (int)l0) - KZG_FP64_2P52;
        d[4 * g + 1] = __hiloint2double(0x43300000, (int)l1) - KZG_FP64_2P52;
        d[4 * g + 2] = __hiloint2double(0x43300000, (int)l2) - KZG_FP64_2P52;
        d[4 * g + 3] = __hiloint2double(0x43300000, (int)l3) - KZG_FP64_2P52;
    }
}

// t[0..15]: un-normalised limbs of a value < 2p  ->  fully reduced 12 x u32
__device__ __forceinline__ Fp fp_from_f64(const double *t) {
    uint32_t l[16];
    double c = 0.0;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        double v = t[j] + c;
        double hi = __dadd_rz(v, KZG_FP64_M76) - KZG_FP64_M76;
        l[j] = (uint32_t)__double2loint((v - hi) + KZG_FP64_2P52);
        c = hi * KZG_FP64_INV24;
    }
    Fp r;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        r.v[3 * g + 0] = l[4 * g] | (l[4 * g + 1] << 24);
        r.v[3 * g + 1] = (l[4 * g + 1] >> 8) | (l[4 * g + 2] << 16);
        r.v[3 * g + 2] = (l[4 * g + 2] >> 16) | (l[4 * g + 3] << 8);
    }
    Fp::final_sub(r.v);
    return r;
}

// one Montgomery row on the accumulator t[0..15] (+ implicit shift): see header
__device__ __forceinline__ void fp64_reduce_row(double *t) {
    double hi = __dadd_rz(t[0], KZG_FP64_M76) - KZG_FP64_M76;
    double lo = t[0] - hi;
    double pr = lo * KZG_FP64_PINV;
    double ph = __dadd_rz(pr, KZG_FP64_M76) - KZG_FP64_M76;
    double q = pr - ph;
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = __fma_rn(q, FP64_P[j], t[j]);
    t[1] = __fma_rn(t[0], KZG_FP64_INV24, t[1]);
#pragma unroll
    for (int j = 0; j < 15; ++j) t[j] = t[j + 1];
    t[15] = 0.0;
}

// a * b * 2^-384 mod p, inputs and output fully reduced 12 x u32 Montgomery operands (same contract as Fp::mul)
__device__ __forceinline__ Fp fp_mul_f64(const Fp &a, const Fp &b) {
    double A[16], B[16], t[16];
    fp_to_f64(A, a.v);
    fp_to_f64(B, b.v);
#pragma unroll
    for (int j = 0; j < 16; ++j) t[j] = A[j] * B[0];
    fp64_reduce_row(t);
#pragma unroll
    for (int i = 1; i < 16; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = __fma_rn(A[j], B[i], t[j]);
        fp64_reduce_row(t);
    }
    return fp_from_f64(t);
}

// a^2 * 2^-384 mod p: the 120 cross products once with a doubled operand (exact), 16 squares
__device__ __forceinline__ Fp fp_sqr_f64(const Fp &a) {
    double A[16], A2[16], t[31];
    fp_to_f64(A, a.v);
#pragma unroll
    for (int j = 0; j < 16; ++j) A2[j] = A[j] + A[j];
#pragma unroll
    for (int k = 0; k < 31; ++k) t[k] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        t[2 * i] = __fma_rn(A[i], A[i], t[2 * i]);
#pragma unroll
        for (int j = i + 1; j < 16; ++j) t[i + j] = __fma_rn(A[i], A2[j], t[i + j]);
    }
    // reduction rows over the 31-column product (column bound: 16 products < 2^48 from a^2 + 16 from q*p + carry < 2^53)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        double hi = __dadd_rz(t[i], KZG_FP64_M76) - KZG_FP64_M76;
        double lo = t[i] - hi;
        double pr = lo * KZG_FP64_PINV;
        double ph = __dadd_rz(pr, KZG_FP64_M76) - KZG_FP64_M76;
        double q = pr - ph;
#pragma unroll
        for (int j = 0; j < 16; ++j) t[i + j] = __fma_rn(q, FP64_P[j], t[i + j]);
        if (i + 1 < 31) t[i + 1] = __fma_rn(t[i], KZG_FP64_INV24, t[i + 1]);
    }
    // t[16..30] hold the result; t[31] would be the carry of column 30, folded by fp_from_f64's chain
    double r[16];
#pragma unroll
    for (int j = 0; j < 15; ++j) r[j] = t[16 + j];
    r[15] = 0.0;
    return fp_from_f64(r);
}

}  // namespace kzg
