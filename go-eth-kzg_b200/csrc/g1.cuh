// G1 group law on device: y^2 = x^3 + 4 over Fp, extended-Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0).
//
// Replaces what the reference gets from gnark-crypto's G1Affine/G1Jac (call sites:
// internal/multiexp/multiexp.go:20-26, internal/domain/fft.go:39,80-83).  All degenerate cases
// (P+P, P+(-P), infinity operands) are handled exactly: spec vectors such as constant blobs and
// verify_cell_kzg_proof_batch "same cell multiple times" do reach them.
//
// Cost in Fp multiplications (mul == sqr here): mixed add 10, full add 14, double 9.
#pragma once
#include "field.cuh"
#include "bingcd.cuh"

namespace kzg {

// Multiplication policies.  MulInline expands the ~600-instruction Montgomery product in place
// (used only in the MSM inner loop, where the ~5% call overhead matters); MulCall goes through
// one shared out-of-line copy (arguments and result stay in registers under the CUDA ABI), which
// keeps every other kernel small enough for the 32 KB instruction cache and quick to compile.
static __device__ __noinline__ Fp fp_mul_ni(Fp a, Fp b) { return Fp::mul(a, b); }
static __device__ __noinline__ Fr fr_mul_ni(Fr a, Fr b) { return Fr::mul(a, b); }
static __device__ __noinline__ Fp fp_sqr_ni(Fp a) { return Fp::sqr(a); }
static __device__ __noinline__ Fr fr_sqr_ni(Fr a) { return Fr::sqr(a); }
static __device__ __noinline__ Fp fp_mam_ni(Fp a, Fp b, Fp c, Fp d) { return Fp::mul_add_mul(a, b, c, d); }
// TWO independent squarings in one out-of-line body: ptxas interleaves their four carry chains, so a lone warp keeps the multiply pipe
// twice as busy (the kernels that use it -- G1 FFT stages, subgroup test -- run with 2-3 warps per scheduler and are latency-limited)
struct FpPair { Fp a, b; };
static __device__ __noinline__ FpPair fp_sqr2_ni(Fp a, Fp b) { FpPair r; r.a = Fp::sqr(a); r.b = Fp::sqr(b); return r; }
// the same for two products, a squaring beside a product, and a one-reduction two-product a*b + c*d beside a product: with these a Jacobian
// doubling is 4 calls deep instead of 7 and a mixed addition 5 instead of 11 (g1fft.cuh: jac_dbl2, jac_madd2)
static __device__ __noinline__ FpPair fp_mul2_ni(Fp a, Fp b, Fp c, Fp d) { FpPair r; r.a = Fp::mul(a, b); r.b = Fp::mul(c, d); return r; }
static __device__ __noinline__ FpPair fp_sqrmul_ni(Fp a, Fp c, Fp d) { FpPair r; r.a = Fp::sqr(a); r.b = Fp::mul(c, d); return r; }
static __device__ __noinline__ FpPair fp_mammul_ni(Fp a, Fp b, Fp c, Fp d, Fp e, Fp f) { FpPair r; r.a = Fp::mul_add_mul(a, b, c, d); r.b = Fp::mul(e, f); return r; }
// msub(a, b, c, d) = a*b - c*d.  The *Lazy policies compute it as a*b + (-c)*d under ONE Montgomery reduction
// (Mont::mul_add_mul: 432 instead of 576 wide-multiply IMADs); the others as two products.
struct MulInline {
    static __device__ __forceinline__ Fp mul(const Fp &a, const Fp &b) { return Fp::mul(a, b); }
    static __device__ __forceinline__ Fp sqr(const Fp &a) { return Fp::sqr(a); }
    static __device__ __forceinline__ Fp msub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return Fp::sub(Fp::mul(a, b), Fp::mul(c, d)); }
    static __device__ __forceinline__ void sqr2(const Fp &a, const Fp &b, Fp &ra, Fp &rb) { ra = Fp::sqr(a); rb = Fp::sqr(b); }
};
struct MulCall {
    static __device__ __forceinline__ Fp mul(const Fp &a, const Fp &b) { return fp_mul_ni(a, b); }
    static __device__ __forceinline__ Fp sqr(const Fp &a) { return fp_sqr_ni(a); }
    static __device__ __forceinline__ Fp msub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return Fp::sub(fp_mul_ni(a, b), fp_mul_ni(c, d)); }
    static __device__ __forceinline__ void sqr2(const Fp &a, const Fp &b, Fp &ra, Fp &rb) { ra = fp_sqr_ni(a); rb = fp_sqr_ni(b); }
};
// sqr2(a, b, ra, rb): ra = a^2, rb = b^2 (independent).  MulCall2 runs them in one dual body (fp_sqr2_ni), the others one after the other.
struct MulInlineLazy : MulInline {
    static __device__ __forceinline__ Fp msub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return Fp::mul_add_mul(a, b, Fp::neg(c), d); }
};
struct MulCallLazy : MulCall {
    static __device__ __forceinline__ Fp msub(const Fp &a, const Fp &b, const Fp &c, const Fp &d) { return fp_mam_ni(a, b, Fp::neg(c), d); }
};
struct MulCall2 : MulCall {
    static __device__ __forceinline__ void sqr2(const Fp &a, const Fp &b, Fp &ra, Fp &rb) { FpPair r = fp_sqr2_ni(a, b); ra = r.a; rb = r.b; }
};

// affine point in Montgomery form; infinity is encoded as (0,0) (not on the curve since b=4)
struct G1Aff {
    Fp x, y;
    __device__ __forceinline__ bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

struct G1 {
    Fp X, Y, ZZ, ZZZ;

    __device__ __forceinline__ bool is_inf() const { return ZZ.is_zero(); }
    static __device__ __forceinline__ G1 infinity() {
        G1 r; r.X = Fp::zero(); r.Y = Fp::zero(); r.ZZ = Fp::zero(); r.ZZZ = Fp::zero();
        return r;
    }
    static __device__ __forceinline__ G1 from_affine(const G1Aff &a) {
        G1 r;
        if (a.is_inf()) return infinity();
        r.X = a.x; r.Y = a.y; r.ZZ = Fp::one(); r.ZZZ = Fp::one();
        return r;
    }
    __device__ __forceinline__ void neg_inplace() { Y = Fp::neg(Y); }
};

// 2*(x,y) from affine (mdbl-2008-s-1).  Cold path: kept out of line.
static __device__ __noinline__ void g1_dbl_affine(G1 *r, const G1Aff *a) {
    typedef MulCall M_;
    Fp U = Fp::dbl(a->y);
    Fp V = M_::sqr(U);
    Fp W = M_::mul(U, V);
    Fp S = M_::mul(a->x, V);
    Fp M = M_::sqr(a->x);
    M = Fp::add(Fp::dbl(M), M);
    Fp X3 = Fp::sub(M_::sqr(M), Fp::dbl(S));
    r->Y = Fp::sub(M_::mul(M, Fp::sub(S, X3)), M_::mul(W, a->y));
    r->X = X3;
    r->ZZ = V;
    r->ZZZ = W;
}

// r = 2*p (dbl-2008-s-1, a = 0)
template <class M_ = MulCall> __device__ __forceinline__ G1 g1_dbl(const G1 &p) {
    if (p.is_inf()) return p;
    G1 r;
    Fp U = Fp::dbl(p.Y);
    Fp V = M_::sqr(U);
    Fp W = M_::mul(U, V);
    Fp S = M_::mul(p.X, V);
    Fp M = M_::sqr(p.X);
    M = Fp::add(Fp::dbl(M), M);
    r.X = Fp::sub(M_::sqr(M), Fp::dbl(S));
    r.Y = Fp::sub(M_::mul(M, Fp::sub(S, r.X)), M_::mul(W, p.Y));
    r.ZZ = M_::mul(V, p.ZZ);
    r.ZZZ = M_::mul(W, p.ZZZ);
    return r;
}
// Value semantics on purpose: g1_dbl_cold is reached from inside other out-of-line functions, and a
// pointer to a register-resident point handed down two levels of cloned callees has been seen to be
// read as a global address by nvcc 12.9 code (compute-sanitizer: invalid __global__ read at a
// local-window address, in g1_add_ool -> g1_dbl_cold).
static __device__ __noinline__ G1 g1_dbl_cold(G1 p) { return g1_dbl<MulCall>(p); }

// acc += (x2,y2)   (madd-2008-s); b must not be the infinity encoding unless checked by caller
template <class M_ = MulCall> __device__ __forceinline__ void g1_add_affine(G1 &acc, const G1Aff &b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc.X = b.x; acc.Y = b.y; acc.ZZ = Fp::one(); acc.ZZZ = Fp::one(); return; }
    Fp U2 = M_::mul(b.x, acc.ZZ);
    Fp S2 = M_::mul(b.y, acc.ZZZ);
    Fp Pd = Fp::sub(U2, acc.X);
    Fp R = Fp::sub(S2, acc.Y);
    if (Pd.is_zero()) {
        if (R.is_zero()) g1_dbl_affine(&acc, &b);
        else acc = G1::infinity();
        return;
    }
    Fp PP = M_::sqr(Pd);
    Fp PPP = M_::mul(Pd, PP);
    Fp Q = M_::mul(acc.X, PP);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(R), PPP), Fp::dbl(Q));
    acc.Y = M_::msub(R, Fp::sub(Q, X3), acc.Y, PPP);
    acc.X = X3;
    acc.ZZ = M_::mul(acc.ZZ, PP);
    acc.ZZZ = M_::mul(acc.ZZZ, PPP);
}

// a += b   (add-2008-s)
template <class M_ = MulCall> __device__ __forceinline__ void g1_add(G1 &a, const G1 &b) {
    if (b.is_inf()) return;
    if (a.is_inf()) { a = b; return; }
    Fp U1 = M_::mul(a.X, b.ZZ);
    Fp U2 = M_::mul(b.X, a.ZZ);
    Fp S1 = M_::mul(a.Y, b.ZZZ);
    Fp S2 = M_::mul(b.Y, a.ZZZ);
    Fp Pd = Fp::sub(U2, U1);
    Fp R = Fp::sub(S2, S1);
    if (Pd.is_zero()) {
        if (R.is_zero()) a = g1_dbl_cold(a);
        else a = G1::infinity();
        return;
    }
    Fp PP = M_::sqr(Pd);
    Fp PPP = M_::mul(Pd, PP);
    Fp Q = M_::mul(U1, PP);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(R), PPP), Fp::dbl(Q));
    a.Y = M_::msub(R, Fp::sub(Q, X3), S1, PPP);
    a.X = X3;
    a.ZZ = M_::mul(M_::mul(a.ZZ, b.ZZ), PP);
    a.ZZZ = M_::mul(M_::mul(a.ZZZ, b.ZZZ), PPP);
}
// out-of-line full add for cold / code-size-sensitive call sites: g1_add_ool for operands in
// shared or global memory, g1_add_v (value semantics) for register-resident operands
static __device__ __noinline__ void g1_add_ool(G1 *a, const G1 *b) { g1_add<MulCallLazy>(*a, *b); }
static __device__ __noinline__ G1 g1_add_v(G1 a, G1 b) { g1_add<MulCallLazy>(a, b); return a; }

// ---- Jacobian coordinates for doubling-heavy code (scalar multiplications, subgroup check) ----
// Twiddle multiplications run in Jacobian coordinates (x = X/Z^2, y = Y/Z^3): a doubling is
// 2M + 5S there against 6M + 3S in XYZZ, and a twiddle multiplication is 132 doublings + ~62
// additions.  Table entries cache Z^2 and Z^3 so an addition is 11M + 3S.
struct G1J { Fp X, Y, Z; };
struct G1JT { Fp X, Y, Z, ZZ, ZZZ; };

template <class M_ = MulCall> __device__ __forceinline__ void jac_dbl(G1J &p) {          // dbl-2009-l, a = 0; no-op on infinity (Z = 0 stays 0)
    Fp A, B, C, T2;
    M_::sqr2(p.X, p.Y, A, B);
    M_::sqr2(B, Fp::add(p.X, B), C, T2);
    Fp D = Fp::dbl(Fp::sub(Fp::sub(T2, A), C));
    Fp E = Fp::add(Fp::dbl(A), A);
    Fp F = M_::sqr(E);
    Fp Z3 = Fp::dbl(M_::mul(p.Y, p.Z));
    p.X = Fp::sub(F, Fp::dbl(D));
    Fp C8 = Fp::dbl(Fp::dbl(Fp::dbl(C)));
    p.Y = Fp::sub(M_::mul(E, Fp::sub(D, p.X)), C8);
    p.Z = Z3;
}
template <class M_ = MulCall> __device__ __forceinline__ void jac_add(G1J &a, const G1JT &b) {   // b is never the point at infinity
    if (a.Z.is_zero()) { a.X = b.X; a.Y = b.Y; a.Z = b.Z; return; }
    Fp Z1Z1 = M_::sqr(a.Z);
    Fp U1 = M_::mul(a.X, b.ZZ), U2 = M_::mul(b.X, Z1Z1);
    Fp S1 = M_::mul(a.Y, b.ZZZ), S2 = M_::mul(M_::mul(b.Y, a.Z), Z1Z1);
    Fp H = Fp::sub(U2, U1), r = Fp::sub(S2, S1);
    if (H.is_zero()) {
        if (r.is_zero()) { a.X = b.X; a.Y = b.Y; a.Z = b.Z; jac_dbl<M_>(a); }
        else a.Z = Fp::zero();
        return;
    }
    Fp HH = M_::sqr(H), HHH = M_::mul(H, HH), V = M_::mul(U1, HH);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(r), HHH), Fp::dbl(V));
    a.Y = M_::msub(r, Fp::sub(V, X3), S1, HHH);
    a.X = X3;
    a.Z = M_::mul(M_::mul(a.Z, b.Z), H);
}
__device__ __forceinline__ G1JT jac_cache(const G1J &p) {
    G1JT t; t.X = p.X; t.Y = p.Y; t.Z = p.Z; t.ZZ = fp_sqr_ni(p.Z); t.ZZZ = fp_mul_ni(t.ZZ, p.Z);
    return t;
}

// XYZZ -> Jacobian with Z = ZZ*ZZZ:  X' = X*ZZ*ZZZ^2,  Y' = Y*ZZ^3*ZZZ^2   (P not at infinity)
__device__ __forceinline__ G1J jac_from_xyzz(const G1 &P) {
    G1J r;
    Fp t = fp_sqr_ni(P.ZZZ);
    r.Z = fp_mul_ni(P.ZZ, P.ZZZ);
    r.X = fp_mul_ni(fp_mul_ni(P.X, P.ZZ), t);
    r.Y = fp_mul_ni(fp_mul_ni(P.Y, fp_mul_ni(fp_sqr_ni(P.ZZ), P.ZZ)), t);
    return r;
}
__device__ __forceinline__ G1 jac_to_xyzz(const G1J &a) {
    G1 r;
    r.X = a.X; r.Y = a.Y; r.ZZ = fp_sqr_ni(a.Z); r.ZZZ = fp_mul_ni(r.ZZ, a.Z);
    return r;
}

// a^(e) for a public, fixed exponent given as plain little-endian 32-bit limbs: fixed 4-bit windows
// (table a^1..a^15 in local memory; the exponent is the same for every thread, so the table index is
// warp-uniform).  For the 379/381-bit exponents used here: ~377 squarings + ~90 products + 14 for the
// table, against ~380 + ~190 for the bit-by-bit ladder.
static __device__ __noinline__ Fp fp_pow(Fp a, const uint32_t *e, int nlimbs) {
    Fp tab[15];
    tab[0] = a;
#pragma unroll 1
    for (int i = 1; i < 15; ++i) tab[i] = (i & 1) ? fp_sqr_ni(tab[i >> 1]) : fp_mul_ni(tab[i - 1], a);   // tab[i] = a^(i+1)
    Fp r = Fp::one();
    bool started = false;
#pragma unroll 1
    for (int w = nlimbs * 8 - 1; w >= 0; --w) {
        if (started) {
#pragma unroll 1
            for (int q = 0; q < 4; ++q) r = fp_sqr_ni(r);
        }
        unsigned d = (e[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) {
            if (started) r = fp_mul_ni(r, tab[d - 1]); else { r = tab[d - 1]; started = true; }
        }
    }
    return r;
}
__device__ __constant__ const uint32_t FP_PM2[12] = {0xffffaaa9u, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                                     0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};   // p-2
// (p+1)/4, for square roots (p = 3 mod 4)
__device__ __constant__ const uint32_t FP_P1D4[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                                      0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};
// a^-1 (0 -> 0) by the binary GCD of bingcd.cuh: the Montgomery operand aR is inverted as a plain integer,
// (aR)^-1 = a^-1 R^-1, and one product with R^3 brings it back to a^-1 R.  ~6x fewer instructions than the
// Fermat chain fp_pow(a, p-2), most of them on the ALU pipe.
static __device__ __noinline__ Fp fp_inv(Fp a) {
    Fp t, r3;
    BinGcd<12>::inverse(t.v, a.v, FP_MOD, FpParams::inv() & 0x7fffffffu);
#pragma unroll
    for (int i = 0; i < 12; ++i) r3.v[i] = FP_R3[i];
    return fp_mul_ni(t, r3);
}

// XYZZ -> affine (one inversion).  x = X * ZZ^2 / ZZZ^2,  y = Y / ZZZ.
__device__ __forceinline__ G1Aff g1_to_affine(const G1 &p) {
    G1Aff r;
    if (p.is_inf()) { r.x = Fp::zero(); r.y = Fp::zero(); return r; }
    Fp i3 = fp_inv(p.ZZZ);
    Fp i2 = fp_mul_ni(fp_sqr_ni(i3), fp_sqr_ni(p.ZZ));   // 1/ZZ
    r.x = fp_mul_ni(p.X, i2);
    r.y = fp_mul_ni(p.Y, i3);
    return r;
}

// ZCash/gnark compressed encoding of an affine point (serialization.go:98-100 -> G1Affine.Bytes)
__device__ __forceinline__ void g1_compress(uint8_t *out48, const G1Aff &a) {
    uint32_t w[12];
    if (a.is_inf()) {
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i] = 0;
        w[11] = 0xc0000000u;
    } else {
        Fp o1 = Fp::zero(); o1.v[0] = 1;
        Fp x = fp_mul_ni(a.x, o1), y = fp_mul_ni(a.y, o1);
        bool largest = !Fp::geq_limbs(FP_HALF, y.v);   // y > (p-1)/2
#pragma unroll
        for (int i = 0; i < 12; ++i) w[i] = x.v[i];
        w[11] |= 0x80000000u | (largest ? 0x20000000u : 0u);
    }
    uint32_t *o = reinterpret_cast<uint32_t *>(out48);   // 48-byte records are 4-byte aligned
#pragma unroll
    for (int i = 0; i < 12; ++i) o[i] = __byte_perm(w[11 - i], 0, 0x0123);   // big-endian
}

}  // namespace kzg
