// Lane-parallel pairing product check, THROUGHPUT form: one check runs on a group of 8 lanes of a warp (namespace pl8).
// pairing_lanes.cuh (namespace pl32) is the latency form, one check per warp; vm_pairing_check picks by batch size.
//
// Replaces gnark-crypto's bls12381.PairingCheck at internal/kzg/kzg_verify.go:88,190 and
// internal/kzg_multi/kzg_verify.go:94 (fixed G2 arguments, lines precomputed in pairing.cuh).
//
// Why: a batch verifier ends in ONE check per verdict, so its latency is on the critical path of
// every call (a whole 4096-blob EIP-4844 batch waits for a single pairing).  Fp12 is used in its
// flat form Fp2[w]/(w^6 - xi), xi = 1+u:  f = sum_{i<6} f_i w^i, and lane i of the group keeps only
// the Fp2 coefficient f_i in registers (tower coefficient map: flat 0,2,4 = c0.c0,c0.c1,c0.c2 and
// flat 1,3,5 = c1.c0,c1.c1,c1.c2, since v = w^2).  Then
//   product:        h_k = sum_{i+j=k} f_i g_j + xi sum_{i+j=k+6} f_i g_j      6 Fp2 products per lane
//   square:         the same sum over unordered pairs                          4 Fp2 products per lane
//   sparse line:    l = a0 + a1 w^2 + w^3:  h_k = f_k a0 + [xi] f_{k-2} a1 + [xi] f_{k-3}   2 products
//   Frobenius:      h_k = conj(f_k) gamma^k                                     1 product, lane-local
//   cyclotomic sqr: Granger-Scott on the Fp4 pairs (f_j, f_{j+3})               1 square + 1 product
// with operands exchanged by warp shuffles (24 words per Fp2).  The critical path of a check drops
// from ~16 k dependent Fp products (one thread) to ~4.6 k.  Lanes 6 and 7 of a group carry zeros.
#pragma once
#include "pairing.cuh"

namespace kzg {
namespace pl8 {

#define KZG_PL8_GROUP 8

struct PL {
    unsigned mask;     // the 8 lanes of this group inside the warp
    int l;             // lane within the group
    int lc;            // min(l, 5): coefficient index used for addressing (idle lanes mirror lane 5's pattern)
};

__device__ __forceinline__ Fp pl_shfl_fp(const PL &c, const Fp &a, int src) {
    Fp r;
#pragma unroll
    for (int k = 0; k < 12; ++k) r.v[k] = __shfl_sync(c.mask, a.v[k], src, KZG_PL8_GROUP);
    return r;
}
__device__ __forceinline__ Fp2 pl_shfl(const PL &c, const Fp2 &a, int src) {
    Fp2 r; r.c0 = pl_shfl_fp(c, a.c0, src); r.c1 = pl_shfl_fp(c, a.c1, src);
    return r;
}
__device__ __forceinline__ Fp fp_select(bool p, const Fp &a, const Fp &b) {
    Fp r;
#pragma unroll
    for (int k = 0; k < 12; ++k) r.v[k] = p ? a.v[k] : b.v[k];
    return r;
}
__device__ __forceinline__ Fp2 fp2_select(bool p, const Fp2 &a, const Fp2 &b) {
    Fp2 r; r.c0 = fp_select(p, a.c0, b.c0); r.c1 = fp_select(p, a.c1, b.c1);
    return r;
}

// h = f * g
static __device__ __noinline__ Fp2 pl_mul(PL c, Fp2 f, Fp2 g) {
    Fp2 s0 = fp2_zero(), s1 = fp2_zero();
#pragma unroll 1
    for (int i = 0; i < 6; ++i) {
        int j = c.lc - i;
        bool wrap = j < 0;
        if (wrap) j += 6;
        Fp2 a = pl_shfl(c, f, i), b = pl_shfl(c, g, j);
        Fp2 p = fp2_mul(a, b);
        s0 = fp2_select(wrap, s0, fp2_add(s0, p));
        s1 = fp2_select(wrap, fp2_add(s1, p), s1);
    }
    return fp2_add(s0, fp2_mul_xi(s1));
}

// h = f^2: per lane 4 slots (i, j, multiplicity, wrapped); slot encoding i | j<<3 | mult<<6 | wrap<<8
__device__ __constant__ const uint16_t PL_SQR_SLOTS[6][4] = {
    {0 | 0 << 3 | 1 << 6 | 0 << 8, 3 | 3 << 3 | 1 << 6 | 1 << 8, 1 | 5 << 3 | 2 << 6 | 1 << 8, 2 | 4 << 3 | 2 << 6 | 1 << 8},
    {0 | 1 << 3 | 2 << 6 | 0 << 8, 2 | 5 << 3 | 2 << 6 | 1 << 8, 3 | 4 << 3 | 2 << 6 | 1 << 8, 0 | 0 << 3 | 0 << 6 | 0 << 8},
    {1 | 1 << 3 | 1 << 6 | 0 << 8, 4 | 4 << 3 | 1 << 6 | 1 << 8, 0 | 2 << 3 | 2 << 6 | 0 << 8, 3 | 5 << 3 | 2 << 6 | 1 << 8},
    {0 | 3 << 3 | 2 << 6 | 0 << 8, 1 | 2 << 3 | 2 << 6 | 0 << 8, 4 | 5 << 3 | 2 << 6 | 1 << 8, 0 | 0 << 3 | 0 << 6 | 0 << 8},
    {2 | 2 << 3 | 1 << 6 | 0 << 8, 5 | 5 << 3 | 1 << 6 | 1 << 8, 0 | 4 << 3 | 2 << 6 | 0 << 8, 1 | 3 << 3 | 2 << 6 | 0 << 8},
    {0 | 5 << 3 | 2 << 6 | 0 << 8, 1 | 4 << 3 | 2 << 6 | 0 << 8, 2 | 3 << 3 | 2 << 6 | 0 << 8, 0 | 0 << 3 | 0 << 6 | 0 << 8}};
static __device__ __noinline__ Fp2 pl_sqr(PL c, Fp2 f) {
    Fp2 s0 = fp2_zero(), s1 = fp2_zero();
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        unsigned e = PL_SQR_SLOTS[c.lc][s];
        int i = e & 7, j = (e >> 3) & 7, mult = (e >> 6) & 3;
        bool wrap = (e >> 8) & 1;
        Fp2 a = pl_shfl(c, f, i), b = pl_shfl(c, f, j);
        Fp2 p = fp2_mul(a, b);
        Fp2 p2 = fp2_dbl(p);
        p = fp2_select(mult == 2, p2, p);
        p = fp2_select(mult == 0, fp2_zero(), p);
        s0 = fp2_select(wrap, s0, fp2_add(s0, p));
        s1 = fp2_select(wrap, fp2_add(s1, p), s1);
    }
    return fp2_add(s0, fp2_mul_xi(s1));
}

// h = f * (a0 + a1 w^2 + w^3)
static __device__ __noinline__ Fp2 pl_mul_by_line(PL c, Fp2 f, Fp2 a0, Fp2 a1) {
    int l = c.lc;
    Fp2 fm2 = pl_shfl(c, f, (l + 4) % 6), fm3 = pl_shfl(c, f, (l + 3) % 6);
    Fp2 t2 = fp2_mul(fm2, a1);
    t2 = fp2_select(l < 2, fp2_mul_xi(t2), t2);
    Fp2 t3 = fp2_select(l < 3, fp2_mul_xi(fm3), fm3);
    return fp2_add(fp2_add(fp2_mul(f, a0), t2), t3);
}

__device__ __forceinline__ Fp2 pl_conj(const PL &c, const Fp2 &f) { return fp2_select(c.l & 1, fp2_neg(f), f); }   // f^(p^6): odd powers of w change sign
__device__ __forceinline__ Fp2 pl_one(const PL &c) { return fp2_select(c.l == 0, fp2_one(), fp2_zero()); }
__device__ __forceinline__ Fp2 pl_frobenius(const PL &c, const Fp2 &f, const Fp2 *g) { return fp2_mul(fp2_conj(f), g[c.lc]); }

// inverse: f^-1 = conj(f) * N^-1 with N = f conj(f) in Fp6 (flat coefficients 0, 2, 4)
static __device__ __noinline__ Fp2 pl_inv(PL c, Fp2 f) {
    Fp2 fc = pl_conj(c, f);
    Fp2 N = pl_mul(c, f, fc);
    Fp2 n0 = pl_shfl(c, N, 0), n1 = pl_shfl(c, N, 2), n2 = pl_shfl(c, N, 4);
    Fp2 t0 = fp2_sub(fp2_sqr(n0), fp2_mul_xi(fp2_mul(n1, n2)));
    Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(n2)), fp2_mul(n0, n1));
    Fp2 t2 = fp2_sub(fp2_sqr(n1), fp2_mul(n0, n2));
    Fp2 d = fp2_inv(fp2_add(fp2_add(fp2_mul(n0, t0), fp2_mul_xi(fp2_mul(n2, t1))), fp2_mul_xi(fp2_mul(n1, t2))));
    Fp2 sel = fp2_select(c.l == 0, t0, fp2_select(c.l == 2, t1, t2));
    Fp2 g = fp2_mul(sel, d);
    g = fp2_select(c.l == 0 || c.l == 2 || c.l == 4, g, fp2_zero());
    return pl_mul(c, fc, g);
}

// Granger-Scott squaring in the cyclotomic subgroup on the pairs A_j = f_j + f_{j+3} s, s = w^3:
//   A^2 = (x^2 + xi y^2) + 2xy s;   out_0 = 3 t0x - 2 f_0, out_3 = 3 t0y + 2 f_3, out_2 = 3 t1x - 2 f_2,
//   out_5 = 3 t1y + 2 f_5, out_1 = 3 xi t2y + 2 f_1, out_4 = 3 t2x - 2 f_4   (same identities as fp12_cyc_sqr)
static __device__ __noinline__ Fp2 pl_cyc_sqr(PL c, Fp2 f) {
    int l = c.lc;
    bool hi = l >= 3;
    Fp2 other = pl_shfl(c, f, hi ? l - 3 : l + 3);
    Fp2 sq = fp2_sqr(f);
    Fp2 xy = fp2_mul(f, other);
    Fp2 osq = pl_shfl(c, sq, hi ? l - 3 : l + 3);
    // x-lane (l < 3): tx = x^2 + xi y^2 ; y-lane: ty = 2xy
    Fp2 t = fp2_select(hi, fp2_dbl(xy), fp2_add(sq, fp2_mul_xi(osq)));
    // routing: lane 0 <- 0, 3 <- 3, 2 <- 1, 5 <- 4, 1 <- 5 (times xi), 4 <- 2
    const int src = l == 0 ? 0 : l == 1 ? 5 : l == 2 ? 1 : l == 3 ? 3 : l == 4 ? 2 : 4;
    Fp2 T = pl_shfl(c, t, src);
    T = fp2_select(l == 1, fp2_mul_xi(T), T);
    Fp2 T3 = fp2_triple(T), f2 = fp2_dbl(f);
    bool plus = (l == 1) || (l == 3) || (l == 5);
    return fp2_select(plus, fp2_add(T3, f2), fp2_sub(T3, f2));
}

// f^|x| then conjugate (x < 0), cyclotomic subgroup only
static __device__ __noinline__ Fp2 pl_pow_x(PL c, Fp2 a) {
    Fp2 acc = a;
#pragma unroll 1
    for (int bit = 62; bit >= 0; --bit) {
        acc = pl_cyc_sqr(c, acc);
        if ((KZG_BLS_X_ABS >> bit) & 1) acc = pl_mul(c, acc, a);
    }
    return pl_conj(c, acc);
}

// final exponentiation, same chain as final_exp() in pairing.cuh
static __device__ __noinline__ Fp2 pl_final_exp(PL c, Fp2 f0, const Fp2 *g) {
    Fp2 f = pl_mul(c, pl_conj(c, f0), pl_inv(c, f0));                        // f^(p^6-1)
    f = pl_mul(c, pl_frobenius(c, pl_frobenius(c, f, g), g), f);             // ^(p^2+1)
    Fp2 fc = pl_conj(c, f);
    Fp2 a = pl_mul(c, pl_pow_x(c, f), fc);                                   // f^(x-1)
    a = pl_mul(c, pl_pow_x(c, a), pl_conj(c, a));                            // f^((x-1)^2)
    Fp2 b = pl_mul(c, pl_pow_x(c, a), pl_frobenius(c, a, g));                // ^(x+p)
    Fp2 u = pl_pow_x(c, pl_pow_x(c, b));                                     // b^(x^2)
    Fp2 bp2 = pl_frobenius(c, pl_frobenius(c, b, g), g);
    Fp2 cc = pl_mul(c, pl_mul(c, u, bp2), pl_conj(c, b));                    // ^(x^2+p^2-1)
    Fp2 f3 = pl_mul(c, pl_cyc_sqr(c, f), f);                                 // f^3
    return pl_mul(c, cc, f3);
}

// multi-Miller loop over (A, Q[qa]) and (B, Q[qb]); A, B in XYZZ (every lane of the group passes the
// same values); an infinite point contributes 1 (gnark's PairingCheck skips points at infinity).
static __device__ __noinline__ Fp2 pl_miller2(PL c, const PairingConsts *pc, const G1 *pA, int qa, const G1 *pB, int qb) {
    // lanes 0 and 1 invert for A and B at the same time: with i = 1/(ZZ*Y): 1/y = ZZZ*ZZ*i, x/y = X*ZZZ*i
    G1 P = (c.l == 1) ? *pB : *pA;
    bool use = !P.is_inf();
    Fp inv = fp_inv(fp_mul_ni(P.ZZ, P.Y));
    Fp py = fp_mul_ni(fp_mul_ni(P.ZZZ, P.ZZ), inv), px = fp_mul_ni(fp_mul_ni(P.X, P.ZZZ), inv);
    // lane q < 4 scales one Fp component of a line: q = 0,1 -> A.c0, A.c1 times 1/y; q = 2,3 -> B.c0, B.c1 times x/y
    Fp sA = fp_select(c.l & 2, pl_shfl_fp(c, px, 0), pl_shfl_fp(c, py, 0));
    Fp sB = fp_select(c.l & 2, pl_shfl_fp(c, px, 1), pl_shfl_fp(c, py, 1));
    const bool useA = __shfl_sync(c.mask, (int)use, 0, KZG_PL8_GROUP) != 0, useB = __shfl_sync(c.mask, (int)use, 1, KZG_PL8_GROUP) != 0;
    const G2Lines *LA = &pc->q[qa], *LB = &pc->q[qb];
    const int comp = c.l & 3;
    Fp2 f = pl_one(c);
    int li = 0;
#pragma unroll 1
    for (int bit = 62; bit >= 0; --bit) {
        f = pl_sqr(c, f);
        int nl = ((KZG_BLS_X_ABS >> bit) & 1) ? 2 : 1;
#pragma unroll 1
        for (int s = 0; s < nl; ++s, ++li) {
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                const G2Lines *L = i ? LB : LA;
                const Fp2 *e = (comp & 2) ? &L->B[li] : &L->A[li];
                const Fp *src = (comp & 1) ? &e->c1 : &e->c0;
                Fp m = fp_mul_ni(*src, i ? sB : sA);
                Fp2 a0, a1;
                a0.c0 = pl_shfl_fp(c, m, 0); a0.c1 = pl_shfl_fp(c, m, 1);
                a1.c0 = pl_shfl_fp(c, m, 2); a1.c1 = pl_shfl_fp(c, m, 3);
                Fp2 h = pl_mul_by_line(c, f, a0, a1);
                f = fp2_select(i ? useB : useA, h, f);
            }
        }
    }
    return pl_conj(c, f);      // x < 0
}

// result[i] = pre_status[i] if that is an error, else OK / VERIFY_FAILED for
//   e(A_i, Q[qa]) * e(B_i, Q[qb]) == 1.     8 lanes per check, 4 checks per warp.
static __global__ void __launch_bounds__(128) k_pairing_lanes(const PairingConsts *__restrict__ pc, const G1 *__restrict__ A, int qa, const G1 *__restrict__ B, int qb,
                                                       const int32_t *pre_status, int32_t *result, size_t n) {   // pre_status may alias result
    const size_t gid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / KZG_PL8_GROUP;
    const size_t warp_first = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) / KZG_PL8_GROUP;
    if (warp_first >= n) return;                       // whole warp idle
    const size_t idx = gid < n ? gid : n - 1;          // padding groups repeat the last check (keeps the warp convergent)
    PL c;
    const int lane = threadIdx.x & 31;
    c.l = lane & (KZG_PL8_GROUP - 1);
    c.lc = c.l < 6 ? c.l : 5;
    c.mask = 0xffu << (lane & ~(KZG_PL8_GROUP - 1));
    G1 a = A[idx], b = B[idx];
    Fp2 f = pl_miller2(c, pc, &a, qa, &b, qb);
    f = fp2_select(c.l < 6, f, fp2_zero());
    Fp2 r = pl_final_exp(c, f, pc->gamma);
    bool ok = c.l == 0 ? fp2_eq(r, fp2_one()) : (c.l < 6 ? fp2_is_zero(r) : true);
    unsigned all = __ballot_sync(c.mask, ok);
    if (gid < n && c.l == 0) {
        int32_t pre = pre_status ? pre_status[idx] : (int32_t)ST_OK;
        result[idx] = pre != ST_OK ? pre : ((all & c.mask) == c.mask ? (int32_t)ST_OK : (int32_t)ST_VERIFY_FAILED);
    }
}

}  // namespace pl8
}  // namespace kzg
