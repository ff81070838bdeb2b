// G2 (the twist y^2 = x^3 + 4(1+u) over Fp2) on device: decompression and the subgroup check of the
// trusted-setup ingestion path.
//
// Replaces gnark-crypto's G2Affine.SetBytes as used by trusted_setup.go:45-83
// (CheckTrustedSetupIsWellFormed: with subgroup check) and trusted_setup.go:124-134 (NoSubgroupChecks).
// Only a few dozen G2 points exist per context, so the code is written for clarity: Jacobian
// coordinates over Fp2, and the subgroup test is the definition, [r]Q == O.
#pragma once
#include "pairing.cuh"

namespace kzg {

struct G2Aff { Fp2 x, y; int inf; };
struct G2Dec { G2Aff a; int32_t st; };      // returned by value (no pointers to registers across out-of-line calls)

// 96 bytes = x.c1 (top three bits: compressed | infinity | y lexicographically largest) || x.c0
static __device__ __noinline__ G2Dec g2_decompress(const uint8_t *p) {
    G2Dec r;
    r.a.x = fp2_zero(); r.a.y = fp2_zero(); r.a.inf = 0; r.st = ST_OK;
    unsigned m = p[0] >> 5;
    if (m != 4 && m != 5 && m != 6) { r.st = ST_BAD_G1_ENCODING; return r; }
    if (m == 6) {
        unsigned o = p[0] & 0x1f;
        for (int i = 1; i < 96; ++i) o |= p[i];
        if (o) r.st = ST_BAD_G1_ENCODING; else r.a.inf = 1;
        return r;
    }
    bool ok = true;
    Fp2 x;
    x.c1 = fp_from_be48(p, true, &ok);
    x.c0 = fp_from_be48(p + 48, false, &ok);
    if (!ok) { r.st = ST_BAD_G1_ENCODING; return r; }
    Fp four = Fp::dbl(Fp::dbl(Fp::one()));
    Fp2 b2; b2.c0 = four; b2.c1 = four;
    Fp2Opt ys = fp2_sqrt(fp2_add(fp2_mul(fp2_sqr(x), x), b2));
    if (!ys.ok) { r.st = ST_NOT_ON_CURVE; return r; }
    Fp2 y = ys.v;
    bool largest = y.c1.is_zero() ? fp_lex_largest(y.c0) : fp_lex_largest(y.c1);
    if (largest != (m == 5)) y = fp2_neg(y);
    r.a.x = x; r.a.y = y;
    return r;
}

struct G2J { Fp2 X, Y, Z; };     // x = X/Z^2, y = Y/Z^3, infinity <=> Z == 0

static __device__ __noinline__ G2J g2j_dbl(G2J p) {                      // dbl-2009-l (a = 0); infinity stays infinity
    Fp2 A = fp2_sqr(p.X), B = fp2_sqr(p.Y), C = fp2_sqr(B);
    Fp2 D = fp2_dbl(fp2_sub(fp2_sub(fp2_sqr(fp2_add(p.X, B)), A), C));
    Fp2 E = fp2_add(fp2_dbl(A), A), F = fp2_sqr(E);
    G2J r;
    r.X = fp2_sub(F, fp2_dbl(D));
    r.Y = fp2_sub(fp2_mul(E, fp2_sub(D, r.X)), fp2_dbl(fp2_dbl(fp2_dbl(C))));
    r.Z = fp2_dbl(fp2_mul(p.Y, p.Z));
    return r;
}
static __device__ __noinline__ G2J g2j_add_affine(G2J a, Fp2 x2, Fp2 y2) {   // madd-2007-bl, all special cases; (x2, y2) finite
    G2J r;
    if (fp2_is_zero(a.Z)) { r.X = x2; r.Y = y2; r.Z = fp2_one(); return r; }
    Fp2 Z1Z1 = fp2_sqr(a.Z);
    Fp2 U2 = fp2_mul(x2, Z1Z1), S2 = fp2_mul(fp2_mul(y2, a.Z), Z1Z1);
    Fp2 H = fp2_sub(U2, a.X), rr = fp2_sub(S2, a.Y);
    if (fp2_is_zero(H)) {
        if (fp2_is_zero(rr)) { G2J q; q.X = x2; q.Y = y2; q.Z = fp2_one(); return g2j_dbl(q); }
        r.X = fp2_zero(); r.Y = fp2_one(); r.Z = fp2_zero();
        return r;
    }
    rr = fp2_dbl(rr);
    Fp2 HH = fp2_sqr(H), I = fp2_dbl(fp2_dbl(HH)), J = fp2_mul(H, I), V = fp2_mul(a.X, I);
    r.X = fp2_sub(fp2_sub(fp2_sqr(rr), J), fp2_dbl(V));
    r.Y = fp2_sub(fp2_mul(rr, fp2_sub(V, r.X)), fp2_dbl(fp2_mul(a.Y, J)));
    r.Z = fp2_sub(fp2_sub(fp2_sqr(fp2_add(a.Z, H)), Z1Z1), HH);
    return r;
}
// [r]Q == O  (r = the scalar-field modulus, 255 bits)
// (inlined into its caller: as an out-of-line function taking the G2Aff aggregate by value it returned wrong
// answers under nvcc 12.9 while the identical ladder written inline was right -- kzgb200_dbg_g2_selftest keeps
// both under test)
__device__ __forceinline__ bool g2_in_subgroup(const G2Aff &q) {
    if (q.inf) return true;
    G2J acc; acc.X = q.x; acc.Y = q.y; acc.Z = fp2_one();        // top bit of r (bit 254)
#pragma unroll 1
    for (int bit = 253; bit >= 0; --bit) {
        acc = g2j_dbl(acc);
        if ((FR_MOD[bit >> 5] >> (bit & 31)) & 1) acc = g2j_add_affine(acc, q.x, q.y);
    }
    return fp2_is_zero(acc.Z);
}

// status[i] of n compressed G2 points: decode, optionally subgroup check
static __global__ void __launch_bounds__(32) k_g2_check(const uint8_t *__restrict__ in96, int32_t *__restrict__ status, size_t n, int subgroup) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G2Dec d = g2_decompress(in96 + i * 96);
    int32_t st = d.st;
    if (st == ST_OK && subgroup && !g2_in_subgroup(d.a)) st = ST_NOT_IN_SUBGROUP;
    status[i] = st;
}

}  // namespace kzg
