// Optimal-ate pairing product check on device, specialised to FIXED G2 arguments.
//
// Replaces gnark-crypto's bls12381.PairingCheck as used at internal/kzg/kzg_verify.go:88,190 and
// internal/kzg_multi/kzg_verify.go:94.  Every G2 point those call sites pass is a trusted-setup
// constant ({G2, [s]G2, [s^64]G2}) once the single-proof check is rewritten by bilinearity
//     e(C - [y]G1, -G2) * e(pi, [s]G2 - [z]G2)  ==  e(C - [y]G1 + [z]pi, -G2) * e(pi, [s]G2),
// so the G2 side of the Miller loop (slopes, line coefficients) is precomputed once per context
// (k_g2_prepare) and a check costs only Fp12 arithmetic plus two Fp scalings per line.
//
// Tower: Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3 - (1+u)), Fp12 = Fp6[w]/(w^2 - v).
// Line through T with slope lambda on the M-type twist, evaluated at P=(xP,yP) and scaled by the
// subfield element w^3/yP (killed by the final exponentiation):
//     l = (lambda*xT - yT)/yP  +  (-lambda * xP/yP) v  +  1 * v w
// One thread per check (the checks of a batch run in parallel; latency per check ~ 20 ms).
#pragma once
#include "codec.cuh"

namespace kzg {

struct Fp2 { Fp c0, c1; };

__device__ __forceinline__ Fp2 fp2_zero() { Fp2 r; r.c0 = Fp::zero(); r.c1 = Fp::zero(); return r; }
__device__ __forceinline__ Fp2 fp2_one() { Fp2 r; r.c0 = Fp::one(); r.c1 = Fp::zero(); return r; }
__device__ __forceinline__ Fp2 fp2_add(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = Fp::add(a.c0, b.c0); r.c1 = Fp::add(a.c1, b.c1); return r; }
__device__ __forceinline__ Fp2 fp2_sub(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = Fp::sub(a.c0, b.c0); r.c1 = Fp::sub(a.c1, b.c1); return r; }
__device__ __forceinline__ Fp2 fp2_neg(const Fp2 &a) { Fp2 r; r.c0 = Fp::neg(a.c0); r.c1 = Fp::neg(a.c1); return r; }
__device__ __forceinline__ Fp2 fp2_dbl(const Fp2 &a) { return fp2_add(a, a); }
__device__ __forceinline__ Fp2 fp2_conj(const Fp2 &a) { Fp2 r; r.c0 = a.c0; r.c1 = Fp::neg(a.c1); return r; }
__device__ __forceinline__ Fp2 fp2_mul_xi(const Fp2 &a) { Fp2 r; r.c0 = Fp::sub(a.c0, a.c1); r.c1 = Fp::add(a.c0, a.c1); return r; }   // * (1+u)
__device__ __forceinline__ bool fp2_is_zero(const Fp2 &a) { return a.c0.is_zero() && a.c1.is_zero(); }
__device__ __forceinline__ bool fp2_eq(const Fp2 &a, const Fp2 &b) { return Fp::eq(a.c0, b.c0) && Fp::eq(a.c1, b.c1); }
static __device__ __noinline__ Fp2 fp2_mul(Fp2 a, Fp2 b) {
    Fp t0 = fp_mul_ni(a.c0, b.c0), t1 = fp_mul_ni(a.c1, b.c1);
    Fp t2 = fp_mul_ni(Fp::add(a.c0, a.c1), Fp::add(b.c0, b.c1));
    Fp2 r; r.c0 = Fp::sub(t0, t1); r.c1 = Fp::sub(Fp::sub(t2, t0), t1);
    return r;
}
static __device__ __noinline__ Fp2 fp2_sqr(Fp2 a) {
    Fp t0 = fp_mul_ni(Fp::add(a.c0, a.c1), Fp::sub(a.c0, a.c1));
    Fp t1 = fp_mul_ni(a.c0, a.c1);
    Fp2 r; r.c0 = t0; r.c1 = Fp::dbl(t1);
    return r;
}
__device__ __forceinline__ Fp2 fp2_mul_fp(const Fp2 &a, const Fp &s) { Fp2 r; r.c0 = fp_mul_ni(a.c0, s); r.c1 = fp_mul_ni(a.c1, s); return r; }
static __device__ __noinline__ Fp2 fp2_inv(Fp2 a) {
    Fp n = fp_inv(Fp::add(fp_sqr_ni(a.c0), fp_sqr_ni(a.c1)));
    Fp2 r; r.c0 = fp_mul_ni(a.c0, n); r.c1 = Fp::neg(fp_mul_ni(a.c1, n));
    return r;
}

struct Fp6 { Fp2 c0, c1, c2; };
struct Fp12 { Fp6 c0, c1; };

static __device__ __noinline__ void fp6_mul(Fp6 *r, const Fp6 *a, const Fp6 *b) {
    Fp2 t0 = fp2_mul(a->c0, b->c0), t1 = fp2_mul(a->c1, b->c1), t2 = fp2_mul(a->c2, b->c2);
    Fp2 r0 = fp2_add(fp2_mul_xi(fp2_sub(fp2_sub(fp2_mul(fp2_add(a->c1, a->c2), fp2_add(b->c1, b->c2)), t1), t2)), t0);
    Fp2 r1 = fp2_add(fp2_sub(fp2_sub(fp2_mul(fp2_add(a->c0, a->c1), fp2_add(b->c0, b->c1)), t0), t1), fp2_mul_xi(t2));
    Fp2 r2 = fp2_add(fp2_sub(fp2_sub(fp2_mul(fp2_add(a->c0, a->c2), fp2_add(b->c0, b->c2)), t0), t2), t1);
    r->c0 = r0; r->c1 = r1; r->c2 = r2;
}
__device__ __forceinline__ void fp6_add(Fp6 *r, const Fp6 *a, const Fp6 *b) { r->c0 = fp2_add(a->c0, b->c0); r->c1 = fp2_add(a->c1, b->c1); r->c2 = fp2_add(a->c2, b->c2); }
__device__ __forceinline__ void fp6_sub(Fp6 *r, const Fp6 *a, const Fp6 *b) { r->c0 = fp2_sub(a->c0, b->c0); r->c1 = fp2_sub(a->c1, b->c1); r->c2 = fp2_sub(a->c2, b->c2); }
__device__ __forceinline__ void fp6_neg(Fp6 *r, const Fp6 *a) { r->c0 = fp2_neg(a->c0); r->c1 = fp2_neg(a->c1); r->c2 = fp2_neg(a->c2); }
__device__ __forceinline__ void fp6_mul_v(Fp6 *r, const Fp6 *a) { Fp2 t = fp2_mul_xi(a->c2); r->c2 = a->c1; r->c1 = a->c0; r->c0 = t; }
static __device__ __noinline__ void fp6_inv(Fp6 *r, const Fp6 *a) {
    Fp2 t0 = fp2_sub(fp2_sqr(a->c0), fp2_mul_xi(fp2_mul(a->c1, a->c2)));
    Fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(a->c2)), fp2_mul(a->c0, a->c1));
    Fp2 t2 = fp2_sub(fp2_sqr(a->c1), fp2_mul(a->c0, a->c2));
    Fp2 d = fp2_inv(fp2_add(fp2_add(fp2_mul(a->c0, t0), fp2_mul_xi(fp2_mul(a->c2, t1))), fp2_mul_xi(fp2_mul(a->c1, t2))));
    r->c0 = fp2_mul(t0, d); r->c1 = fp2_mul(t1, d); r->c2 = fp2_mul(t2, d);
}

static __device__ __noinline__ void fp12_mul(Fp12 *r, const Fp12 *a, const Fp12 *b) {
    Fp6 t0, t1, s0, s1, m;
    fp6_mul(&t0, &a->c0, &b->c0);
    fp6_mul(&t1, &a->c1, &b->c1);
    fp6_add(&s0, &a->c0, &a->c1);
    fp6_add(&s1, &b->c0, &b->c1);
    fp6_mul(&m, &s0, &s1);
    fp6_sub(&m, &m, &t0);
    fp6_sub(&m, &m, &t1);
    Fp6 t1v; fp6_mul_v(&t1v, &t1);
    fp6_add(&r->c0, &t0, &t1v);
    r->c1 = m;
}
// complex squaring: (c0 + c1 w)^2 = (c0 + c1)(c0 + v c1) - t - v t  +  2 t w,  t = c0 c1   (2 Fp6 products)
static __device__ __noinline__ void fp12_sqr(Fp12 *r, const Fp12 *a) {
    Fp6 t, s0, s1, m, tv;
    fp6_mul(&t, &a->c0, &a->c1);
    fp6_add(&s0, &a->c0, &a->c1);
    fp6_mul_v(&s1, &a->c1);
    fp6_add(&s1, &s1, &a->c0);
    fp6_mul(&m, &s0, &s1);
    fp6_mul_v(&tv, &t);
    fp6_sub(&m, &m, &t);
    fp6_sub(&r->c0, &m, &tv);
    fp6_add(&r->c1, &t, &t);
}
// g * (a0 + a1 v) for g in Fp6 (6 Fp2 products)
static __device__ __noinline__ void fp6_mul_by_01(Fp6 *r, const Fp6 *g, Fp2 a0, Fp2 a1) {
    Fp2 r0 = fp2_add(fp2_mul(g->c0, a0), fp2_mul_xi(fp2_mul(g->c2, a1)));
    Fp2 r1 = fp2_add(fp2_mul(g->c0, a1), fp2_mul(g->c1, a0));
    Fp2 r2 = fp2_add(fp2_mul(g->c1, a1), fp2_mul(g->c2, a0));
    r->c0 = r0; r->c1 = r1; r->c2 = r2;
}
// f *= l for the sparse line l = (a0 + a1 v) + (v) w:
//   f l = (f0 L0 + v^2 f1) + (v f0 + f1 L0) w,   L0 = a0 + a1 v        (12 Fp2 products)
static __device__ __noinline__ void fp12_mul_by_line(Fp12 *f, Fp2 a0, Fp2 a1) {
    Fp6 x, y, t, u;
    fp6_mul_by_01(&x, &f->c0, a0, a1);
    fp6_mul_by_01(&y, &f->c1, a0, a1);
    fp6_mul_v(&t, &f->c1); fp6_mul_v(&u, &t);          // v^2 f1
    Fp6 vf0; fp6_mul_v(&vf0, &f->c0);
    fp6_add(&f->c0, &x, &u);
    fp6_add(&f->c1, &vf0, &y);
}
__device__ __forceinline__ void fp12_conj(Fp12 *r, const Fp12 *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
__device__ __forceinline__ void fp12_one(Fp12 *r) {
    r->c0.c0 = fp2_one(); r->c0.c1 = fp2_zero(); r->c0.c2 = fp2_zero();
    r->c1.c0 = fp2_zero(); r->c1.c1 = fp2_zero(); r->c1.c2 = fp2_zero();
}
static __device__ __noinline__ void fp12_inv(Fp12 *r, const Fp12 *a) {
    Fp6 t0, t1, d;
    fp6_mul(&t0, &a->c0, &a->c0);
    fp6_mul(&t1, &a->c1, &a->c1);
    Fp6 t1v; fp6_mul_v(&t1v, &t1);
    fp6_sub(&t0, &t0, &t1v);
    fp6_inv(&d, &t0);
    fp6_mul(&r->c0, &a->c0, &d);
    Fp6 m; fp6_mul(&m, &a->c1, &d);
    fp6_neg(&r->c1, &m);
}
__device__ __forceinline__ bool fp12_is_one(const Fp12 *a) {
    return fp2_eq(a->c0.c0, fp2_one()) && fp2_is_zero(a->c0.c1) && fp2_is_zero(a->c0.c2) && fp2_is_zero(a->c1.c0) &&
           fp2_is_zero(a->c1.c1) && fp2_is_zero(a->c1.c2);
}

// Precomputed per context: Frobenius coefficients gamma^k, gamma = (1+u)^((p-1)/6), and for each
// fixed G2 point the 68 line coefficient pairs (A = lambda*xT - yT, B = -lambda).
#define KZG_N_LINES 68
struct G2Lines { Fp2 A[KZG_N_LINES], B[KZG_N_LINES]; };
struct PairingConsts {
    Fp2 gamma[6];
    G2Lines q[3];          // 0: G2 generator, 1: [s]G2, 2: [s^64]G2
};

__device__ __forceinline__ void fp12_frobenius(Fp12 *r, const Fp12 *a, const Fp2 *g) {
    r->c0.c0 = fp2_conj(a->c0.c0);
    r->c0.c1 = fp2_mul(fp2_conj(a->c0.c1), g[2]);
    r->c0.c2 = fp2_mul(fp2_conj(a->c0.c2), g[4]);
    r->c1.c0 = fp2_mul(fp2_conj(a->c1.c0), g[1]);
    r->c1.c1 = fp2_mul(fp2_conj(a->c1.c1), g[3]);
    r->c1.c2 = fp2_mul(fp2_conj(a->c1.c2), g[5]);
}

#define KZG_BLS_X_ABS 0xd201000000010000ULL

// Granger-Scott squaring for elements of the cyclotomic subgroup (everything after the easy part of
// the final exponentiation).  View Fp12 = Fp4[w]/(w^3 - s), Fp4 = Fp2[s]/(s^2 - xi), s = v w:
//   a = A0 + A1 w + A2 w^2,  A0 = g0 + h1 s,  A1 = h0 + g2 s,  A2 = g1 + h2 s
//   a^2 = (3 A0^2 - 2 conj A0) + (3 s A2^2 + 2 conj A1) w + (3 A1^2 - 2 conj A2) w^2
// 9 Fp2 squarings = 18 Fp products (validated against the generic square in the oracle:
// ko_dbg_cyclotomic_sqr_check).
__device__ __forceinline__ void fp4_sqr(Fp2 &rx, Fp2 &ry, const Fp2 &x, const Fp2 &y) {
    Fp2 xx = fp2_sqr(x), yy = fp2_sqr(y);
    rx = fp2_add(xx, fp2_mul_xi(yy));
    ry = fp2_sub(fp2_sub(fp2_sqr(fp2_add(x, y)), xx), yy);
}
__device__ __forceinline__ Fp2 fp2_triple(const Fp2 &z) { return fp2_add(fp2_dbl(z), z); }
static __device__ __noinline__ void fp12_cyc_sqr(Fp12 *r, const Fp12 *a) {
    Fp2 x0 = a->c0.c0, y0 = a->c1.c1, x1 = a->c1.c0, y1 = a->c0.c2, x2 = a->c0.c1, y2 = a->c1.c2;
    Fp2 t0x, t0y, t1x, t1y, t2x, t2y;
    fp4_sqr(t0x, t0y, x0, y0);
    fp4_sqr(t1x, t1y, x1, y1);
    fp4_sqr(t2x, t2y, x2, y2);
    r->c0.c0 = fp2_sub(fp2_triple(t0x), fp2_dbl(x0));
    r->c1.c1 = fp2_add(fp2_triple(t0y), fp2_dbl(y0));
    r->c1.c0 = fp2_add(fp2_triple(fp2_mul_xi(t2y)), fp2_dbl(x1));
    r->c0.c2 = fp2_sub(fp2_triple(t2x), fp2_dbl(y1));
    r->c0.c1 = fp2_sub(fp2_triple(t1x), fp2_dbl(x2));
    r->c1.c2 = fp2_add(fp2_triple(t1y), fp2_dbl(y2));
}

// f^|x| then conjugate (x < 0); valid in the cyclotomic subgroup where inverse == conjugate
static __device__ __noinline__ void fp12_pow_x(Fp12 *r, const Fp12 *a) {
    Fp12 acc = *a;
#pragma unroll 1
    for (int bit = 62; bit >= 0; --bit) {
        fp12_cyc_sqr(&acc, &acc);
        if ((KZG_BLS_X_ABS >> bit) & 1) { Fp12 t = acc; fp12_mul(&acc, &t, a); }
    }
    fp12_conj(r, &acc);
}

// final exponentiation; easy part then the hard part with exponent
// 3*(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3   (Hayashida-Hayasaka-Teruya)
static __device__ __noinline__ void final_exp(Fp12 *r, const Fp12 *f0, const Fp2 *g) {
    Fp12 f, t, u;
    fp12_conj(&t, f0);
    fp12_inv(&u, f0);
    fp12_mul(&f, &t, &u);                     // f^(p^6-1)
    fp12_frobenius(&t, &f, g);
    fp12_frobenius(&u, &t, g);
    t = f; fp12_mul(&f, &u, &t);              // ^(p^2+1)
    Fp12 a, b, c, fc;
    fp12_conj(&fc, &f);
    fp12_pow_x(&t, &f); fp12_mul(&a, &t, &fc);            // f^(x-1)
    fp12_conj(&u, &a);
    fp12_pow_x(&t, &a); { Fp12 t2 = t; fp12_mul(&a, &t2, &u); }   // f^((x-1)^2)
    fp12_pow_x(&t, &a); fp12_frobenius(&u, &a, g); fp12_mul(&b, &t, &u);   // ^(x+p)
    fp12_pow_x(&t, &b); fp12_pow_x(&u, &t);                               // b^(x^2)
    Fp12 bp2, bc;
    fp12_frobenius(&t, &b, g); fp12_frobenius(&bp2, &t, g);
    fp12_conj(&bc, &b);
    fp12_mul(&t, &u, &bp2); fp12_mul(&c, &t, &bc);                        // ^(x^2+p^2-1)
    fp12_cyc_sqr(&t, &f); fp12_mul(&u, &t, &f);                           // f^3
    fp12_mul(r, &c, &u);
}

// multi-Miller loop over two (P_i, fixed Q_i) pairs sharing the squarings.  P given affine; an
// infinite P contributes 1 (gnark's PairingCheck skips points at infinity).
static __device__ __noinline__ void miller2(Fp12 *f, const G1Aff *P0, const G2Lines *L0, const G1Aff *P1, const G2Lines *L1) {
    Fp px[2], py[2];
    bool use[2];
    const G1Aff *Ps[2] = {P0, P1};
    const G2Lines *Ls[2] = {L0, L1};
    for (int i = 0; i < 2; ++i) {
        use[i] = !Ps[i]->is_inf();
        if (use[i]) { py[i] = fp_inv(Ps[i]->y); px[i] = fp_mul_ni(Ps[i]->x, py[i]); }
    }
    fp12_one(f);
    int li = 0;
#pragma unroll 1
    for (int bit = 62; bit >= 0; --bit) {
        fp12_sqr(f, f);
        int nl = ((KZG_BLS_X_ABS >> bit) & 1) ? 2 : 1;
#pragma unroll 1
        for (int s = 0; s < nl; ++s, ++li) {
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
                if (!use[i]) continue;
                fp12_mul_by_line(f, fp2_mul_fp(Ls[i]->A[li], py[i]), fp2_mul_fp(Ls[i]->B[li], px[i]));
            }
        }
    }
    Fp12 t = *f;
    fp12_conj(f, &t);      // x < 0
}

// ---- context init: G2 decompression + line precomputation (one thread per fixed G2 point) --------
// Square roots return VALUES (result + flag): see DESIGN.md section 7 on pointers to registers across
// out-of-line calls.
struct FpOpt { Fp v; bool ok; };
struct Fp2Opt { Fp2 v; bool ok; };
static __device__ __noinline__ FpOpt fp_sqrt_checked(Fp a) {
    FpOpt r;
    r.v = fp_pow(a, FP_P1D4, 12);
    r.ok = Fp::eq(fp_sqr_ni(r.v), a);
    return r;
}
__device__ __forceinline__ bool fp_lex_largest(const Fp &a) {
    Fp o1 = Fp::zero(); o1.v[0] = 1;
    Fp p = fp_mul_ni(a, o1);
    return !Fp::geq_limbs(FP_HALF, p.v);
}
// sqrt in Fp2 by the norm method (result verified by squaring)
static __device__ __noinline__ Fp2Opt fp2_sqrt(Fp2 a) {
    Fp2Opt out; out.v = a; out.ok = true;
    if (fp2_is_zero(a)) return out;
    out.ok = false;
    Fp two = Fp::dbl(Fp::one());
    Fp inv2 = fp_inv(two);
    if (a.c1.is_zero()) {
        FpOpt s = fp_sqrt_checked(a.c0);
        if (s.ok) { out.v.c0 = s.v; out.v.c1 = Fp::zero(); out.ok = true; return out; }
        s = fp_sqrt_checked(Fp::neg(a.c0));
        if (s.ok) { out.v.c0 = Fp::zero(); out.v.c1 = s.v; out.ok = true; }
        return out;
    }
    Fp n = Fp::add(fp_sqr_ni(a.c0), fp_sqr_ni(a.c1));
    FpOpt s = fp_sqrt_checked(n);
    if (!s.ok) return out;
    FpOpt x0 = fp_sqrt_checked(fp_mul_ni(Fp::add(a.c0, s.v), inv2));
    if (!x0.ok) {
        x0 = fp_sqrt_checked(fp_mul_ni(Fp::sub(a.c0, s.v), inv2));
        if (!x0.ok) return out;
    }
    Fp2 r; r.c0 = x0.v; r.c1 = fp_mul_ni(a.c1, fp_inv(Fp::dbl(x0.v)));
    if (!fp2_eq(fp2_sqr(r), a)) return out;
    out.v = r; out.ok = true;
    return out;
}
__device__ __forceinline__ Fp fp_from_be48(const uint8_t *p, bool mask_flags, bool *ok) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p);
    Fp x;
#pragma unroll
    for (int i = 0; i < 12; ++i) x.v[11 - i] = __byte_perm(w[i], 0, 0x0123);
    if (mask_flags) x.v[11] &= 0x1fffffffu;
    if (Fp::geq_limbs(x.v, FP_MOD)) *ok = false;
    Fp r2;
#pragma unroll
    for (int i = 0; i < 12; ++i) r2.v[i] = FP_R2[i];
    return fp_mul_ni(x, r2);
}
// in96[3][96]: compressed G2 points (x.c1 with flags || x.c0), no subgroup check
// (trusted_setup.go:130-133 uses NoSubgroupChecks).  Writes gamma and the three line tables.
static __global__ void k_g2_prepare(const uint8_t *__restrict__ in96, PairingConsts *__restrict__ pc, int32_t *__restrict__ bad) {
    const int qi = threadIdx.x;
    if (qi == 3) {
        // gamma = (1+u)^((p-1)/6)
        const uint32_t E[12] = {0xfffff1c7u, 0x49aa7fffu, 0x72e35555u, 0x051caaaau, 0xd3c82906u, 0xe688231au, 0x7deb831fu, 0xe613e1ebu, 0xb5e1f223u, 0x0c849bf3u, 0x5eeaa66fu, 0x045582fcu};
        Fp2 xi; xi.c0 = Fp::one(); xi.c1 = Fp::one();
        Fp2 g = fp2_one();
        for (int i = 12 * 32 - 1; i >= 0; --i) {
            g = fp2_sqr(g);
            if ((E[i >> 5] >> (i & 31)) & 1) g = fp2_mul(g, xi);
        }
        pc->gamma[0] = fp2_one();
        for (int i = 1; i < 6; ++i) pc->gamma[i] = fp2_mul(pc->gamma[i - 1], g);
        return;
    }
    if (qi > 3) return;
    const uint8_t *p = in96 + qi * 96;
    unsigned m = p[0] >> 5;
    if (m != 4 && m != 5) { atomicMax(bad, (int32_t)ST_BAD_G1_ENCODING); return; }   // infinity is not a usable SRS point
    bool ok = true;
    Fp2 x;
    x.c1 = fp_from_be48(p, true, &ok);
    x.c0 = fp_from_be48(p + 48, false, &ok);
    if (!ok) { atomicMax(bad, (int32_t)ST_BAD_G1_ENCODING); return; }
    Fp four = Fp::dbl(Fp::dbl(Fp::one()));
    Fp2 b2; b2.c0 = four; b2.c1 = four;
    Fp2Opt ys = fp2_sqrt(fp2_add(fp2_mul(fp2_sqr(x), x), b2));
    if (!ys.ok) { atomicMax(bad, (int32_t)ST_NOT_ON_CURVE); return; }
    Fp2 y = ys.v;
    bool largest = y.c1.is_zero() ? fp_lex_largest(y.c0) : fp_lex_largest(y.c1);
    if (largest != (m == 5)) y = fp2_neg(y);
    // affine Miller-loop walk over Q, recording the line coefficients
    Fp2 tx = x, ty = y;
    G2Lines *L = &pc->q[qi];
    int li = 0;
    for (int bit = 62; bit >= 0; --bit) {
        Fp2 xx = fp2_sqr(tx);
        Fp2 lam = fp2_mul(fp2_add(fp2_dbl(xx), xx), fp2_inv(fp2_dbl(ty)));
        L->A[li] = fp2_sub(fp2_mul(lam, tx), ty);
        L->B[li] = fp2_neg(lam);
        ++li;
        Fp2 x3 = fp2_sub(fp2_sqr(lam), fp2_dbl(tx));
        Fp2 y3 = fp2_sub(fp2_mul(lam, fp2_sub(tx, x3)), ty);
        tx = x3; ty = y3;
        if ((KZG_BLS_X_ABS >> bit) & 1) {
            Fp2 lam2 = fp2_mul(fp2_sub(y, ty), fp2_inv(fp2_sub(x, tx)));
            L->A[li] = fp2_sub(fp2_mul(lam2, tx), ty);
            L->B[li] = fp2_neg(lam2);
            ++li;
            Fp2 x4 = fp2_sub(fp2_sub(fp2_sqr(lam2), tx), x);
            Fp2 y4 = fp2_sub(fp2_mul(lam2, fp2_sub(tx, x4)), ty);
            tx = x4; ty = y4;
        }
    }
}

// prod_i e(P_i, Q_{qsel_i}) == 1 ?  for two pairs
static __device__ __noinline__ bool pairing_check2(const PairingConsts *pc, const G1Aff *P0, int q0, const G1Aff *P1, int q1) {
    Fp12 f, r;
    miller2(&f, P0, &pc->q[q0], P1, &pc->q[q1]);
    final_exp(&r, &f, pc->gamma);
    return fp12_is_one(&r);
}

}  // namespace kzg
