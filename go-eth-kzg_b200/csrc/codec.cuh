// Byte-level (de)serialisation on device: big-endian 32-byte scalars and ZCash-format 48-byte
// compressed G1 points.  Mirrors serialization.go:98-216 (which defers to gnark-crypto's
// SetBytesCanonical / G1Affine.SetBytes / Bytes).
#pragma once
#include "g1.cuh"

namespace kzg {

enum : int32_t {
    ST_OK = 0, ST_VERIFY_FAILED = 1, ST_NON_CANONICAL_SCALAR = 2, ST_BAD_G1_ENCODING = 3, ST_NOT_ON_CURVE = 4,
    ST_NOT_IN_SUBGROUP = 5, ST_LENGTH_MISMATCH = 6, ST_BAD_CELL_INDEX = 7, ST_CELL_IDS_NOT_ASCENDING = 8,
    ST_NOT_ENOUGH_CELLS = 9, ST_BAD_ROW_INDEX = 10, ST_ERR_ARGS = 11, ST_ERR_SETUP = 12, ST_ERR_CUDA = 100
};

// 32 big-endian bytes (4-byte aligned) -> 8 little-endian limbs (plain integer)
__device__ __forceinline__ void load_be32(uint32_t *limbs, const uint8_t *p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) limbs[7 - i] = __byte_perm(w[i], 0, 0x0123);
}
__device__ __forceinline__ void store_be32(uint8_t *p, const uint32_t *limbs) {
    uint32_t *w = reinterpret_cast<uint32_t *>(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = __byte_perm(limbs[7 - i], 0, 0x0123);
}
// canonical check: value < r   (serialization.go:141 SetBytesCanonical)
__device__ __forceinline__ bool fr_is_canonical(const uint32_t *limbs) { return !Fr::geq_limbs(limbs, FR_MOD); }

// 48 compressed bytes -> affine point (Montgomery).  gnark-crypto G1Affine.SetBytes semantics:
// top 3 bits = compressed | infinity | y-lexicographically-largest.  No subgroup check here.
__device__ __forceinline__ int32_t g1_decompress(G1Aff &out, const uint8_t *in48) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(in48);
    uint32_t l[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) l[11 - i] = __byte_perm(w[i], 0, 0x0123);
    uint32_t m = l[11] >> 29;
    l[11] &= 0x1fffffffu;
    out.x = Fp::zero(); out.y = Fp::zero();
    if (m != 4 && m != 5 && m != 6) return ST_BAD_G1_ENCODING;   // uncompressed forms need 96 bytes
    if (m == 6) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 12; ++i) o |= l[i];
        return o ? ST_BAD_G1_ENCODING : ST_OK;
    }
    if (Fp::geq_limbs(l, FP_MOD)) return ST_BAD_G1_ENCODING;
    Fp x;
#pragma unroll
    for (int i = 0; i < 12; ++i) x.v[i] = l[i];
    {
        Fp r2;
#pragma unroll
        for (int i = 0; i < 12; ++i) r2.v[i] = FP_R2[i];
        x = fp_mul_ni(x, r2);
    }
    Fp four = Fp::one(); four = Fp::dbl(Fp::dbl(four));
    Fp y2 = Fp::add(fp_mul_ni(fp_sqr_ni(x), x), four);
    Fp y = fp_pow(y2, FP_P1D4, 12);
    if (!Fp::eq(fp_sqr_ni(y), y2)) return ST_NOT_ON_CURVE;
    Fp o1 = Fp::zero(); o1.v[0] = 1;
    Fp yp = fp_mul_ni(y, o1);
    bool largest = !Fp::geq_limbs(FP_HALF, yp.v);
    if (largest != (m == 5)) y = Fp::neg(y);
    out.x = x; out.y = y;
    return ST_OK;
}

// Subgroup membership of an on-curve affine point (the check gnark's SetBytes performs,
// serialization.go:108-115).  Endomorphism test (M. Scott, "A note on group membership tests for
// G1, G2 and GT on BLS pairing-friendly curves", 2021): P is in G1  <=>  phi2(P) == [-x^2]P where
// phi2(x,y) = (beta^2 x, y) and x = -0xd201000000010000.  Cost: 126 doublings + 10 additions.
template <class M_ = MulCall> static __device__ __noinline__ bool g1_in_subgroup(const G1Aff *pa, const uint32_t *beta2) {
    G1Aff a = *pa;
    if (a.is_inf()) return true;
    G1J q; q.X = a.x; q.Y = a.y; q.Z = Fp::one();
    const unsigned long long X_ABS = 0xd201000000010000ULL;
#pragma unroll 1
    for (int rep = 0; rep < 2; ++rep) {               // q = [|x|] q, twice
        G1JT base = jac_cache(q);
#pragma unroll 1
        for (int bit = 62; bit >= 0; --bit) {          // |x| MSB first after the leading 1
            jac_dbl<M_>(q);
            if ((X_ABS >> bit) & 1) jac_add<M_>(q, base);
        }
    }
    // need q == -phi2(P) = (beta2*x, -y):  X == beta2*x*Z^2,  Y == -y*Z^3
    if (q.Z.is_zero()) return false;
    Fp b2;
#pragma unroll
    for (int i = 0; i < 12; ++i) b2.v[i] = beta2[i];
    Fp zz = fp_sqr_ni(q.Z);
    Fp ex = fp_mul_ni(fp_mul_ni(b2, a.x), zz);
    Fp ey = fp_mul_ni(Fp::neg(a.y), fp_mul_ni(zz, q.Z));
    return Fp::eq(ex, q.X) && Fp::eq(ey, q.Y);
}

}  // namespace kzg
