// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
//
// CPU oracle, part 1: BLS12-381 arithmetic (Fp, Fr, Fp2/6/12, G1, G2, optimal-ate
// pairing check, ZCash-format point codec).
//
// The reference (go-eth-kzg) delegates all of this to the un-vendored third-party module
//     github.com/consensys/gnark-crypto v0.16.0   (go.mod:8)
// which is NOT under /root/reference and cannot be built here (no Go toolchain).  This file
// restates the published BLS12-381 algorithms that module implements; it is anchored to the
// reference through its call sites (serialization.go:98-115,134-164, internal/multiexp/
// multiexp.go:20-26, internal/kzg/kzg_verify.go:88,190, internal/kzg_multi/kzg_verify.go:94)
// and pinned by the reference's 311 consensus-spec vectors + KATs (tests/test_oracle_*.py).
//
// Independent of the CUDA product code on purpose: 64-bit limbs + unsigned __int128 here,
// 32-bit limbs + PTX carry chains there; Jacobian coordinates here, XYZZ there.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <stdexcept>

namespace ko {

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------
// generic little-endian multi-limb helpers
// ---------------------------------------------------------------------------------------
template <int N> static inline int cmp_limbs(const u64 *a, const u64 *b) {
    for (int i = N - 1; i >= 0; --i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
template <int N> static inline u64 add_limbs(u64 *r, const u64 *a, const u64 *b) {
    u128 c = 0;
    for (int i = 0; i < N; ++i) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
template <int N> static inline u64 sub_limbs(u64 *r, const u64 *a, const u64 *b) {
    u64 borrow = 0;
    for (int i = 0; i < N; ++i) {
        u128 d = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)d; borrow = (u64)(d >> 64) & 1;
    }
    return borrow;
}

// ---------------------------------------------------------------------------------------
// Montgomery prime field with N 64-bit limbs.  Tag distinguishes Fp / Fr statics.
// ---------------------------------------------------------------------------------------
template <int N, int Tag> struct Field {
    u64 v[N];   // Montgomery form, always fully reduced (< modulus)

    static u64 MOD[N];
    static u64 INV;          // -MOD^{-1} mod 2^64
    static Field R1, R2;     // R mod p, R^2 mod p  (R = 2^(64N))
    static u64 PM2[N];       // p-2
    static u64 HALF[N];      // (p-1)/2

    static void init(const u64 *mod) {
        memcpy(MOD, mod, sizeof(MOD));
        u64 x = 1;                     // Newton: x = mod^{-1} mod 2^64
        for (int i = 0; i < 6; ++i) x *= 2 - MOD[0] * x;
        INV = (u64)0 - x;
        // R1 = 2^(64N) mod p by 64N doublings of 1; R2 by 64N more.
        u64 t[N] = {1};
        for (int i = 0; i < 2 * 64 * N; ++i) {
            u64 c = add_limbs<N>(t, t, t);
            if (c || cmp_limbs<N>(t, MOD) >= 0) sub_limbs<N>(t, t, MOD);
            if (i == 64 * N - 1) memcpy(R1.v, t, sizeof(t));
        }
        memcpy(R2.v, t, sizeof(t));
        u64 two[N] = {2}, one[N] = {1};
        sub_limbs<N>(PM2, MOD, two);
        sub_limbs<N>(HALF, MOD, one);
        for (int i = 0; i < N; ++i) HALF[i] = (HALF[i] >> 1) | (i + 1 < N ? HALF[i + 1] << 63 : 0);
    }

    static Field zero() { Field z; memset(z.v, 0, sizeof(z.v)); return z; }
    static Field one() { return R1; }
    static Field from_u64(u64 x) { Field a = zero(); a.v[0] = x; return a * R2; }
    // plain (non-Montgomery) little-endian limbs, must be < modulus
    static Field from_limbs(const u64 *l) { Field a; memcpy(a.v, l, sizeof(a.v)); return a * R2; }
    void to_limbs(u64 *l) const {
        Field o = zero(); o.v[0] = 1;
        Field t = (*this) * o;
        memcpy(l, t.v, sizeof(t.v));
    }
    // big-endian bytes (8N of them). Returns false if value >= modulus (non canonical).
    static bool from_bytes_be(Field &out, const uint8_t *b) {
        u64 l[N];
        for (int i = 0; i < N; ++i) {
            u64 w = 0;
            for (int j = 0; j < 8; ++j) w = (w << 8) | b[(N - 1 - i) * 8 + j];
            l[i] = w;
        }
        if (cmp_limbs<N>(l, MOD) >= 0) return false;
        out = from_limbs(l);
        return true;
    }
    // big-endian bytes reduced mod p (gnark fr.Element.SetBytes semantics for 8N-byte input)
    static Field from_bytes_be_reduce(const uint8_t *b) {
        u64 l[N];
        for (int i = 0; i < N; ++i) {
            u64 w = 0;
            for (int j = 0; j < 8; ++j) w = (w << 8) | b[(N - 1 - i) * 8 + j];
            l[i] = w;
        }
        while (cmp_limbs<N>(l, MOD) >= 0) sub_limbs<N>(l, l, MOD);
        return from_limbs(l);
    }
    void to_bytes_be(uint8_t *b) const {
        u64 l[N]; to_limbs(l);
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < 8; ++j) b[(N - 1 - i) * 8 + j] = (uint8_t)(l[i] >> (56 - 8 * j));
    }

    bool is_zero() const { u64 o = 0; for (int i = 0; i < N; ++i) o |= v[i]; return o == 0; }
    bool operator==(const Field &b) const { return memcmp(v, b.v, sizeof(v)) == 0; }
    bool operator!=(const Field &b) const { return !(*this == b); }
    // value > (p-1)/2 ?  ("lexicographically largest" in the ZCash encoding)
    bool lex_largest() const { u64 l[N]; to_limbs(l); return cmp_limbs<N>(l, HALF) > 0; }

    Field operator+(const Field &b) const {
        Field r; u64 c = add_limbs<N>(r.v, v, b.v);
        if (c || cmp_limbs<N>(r.v, MOD) >= 0) sub_limbs<N>(r.v, r.v, MOD);
        return r;
    }
    Field operator-(const Field &b) const {
        Field r; if (sub_limbs<N>(r.v, v, b.v)) add_limbs<N>(r.v, r.v, MOD);
        return r;
    }
    Field neg() const { if (is_zero()) return *this; Field r; sub_limbs<N>(r.v, MOD, v); return r; }
    Field dbl() const { return *this + *this; }

    // CIOS Montgomery product
    Field operator*(const Field &b) const {
        u64 t[N + 2];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < N; ++i) {
            u128 c = 0;
            for (int j = 0; j < N; ++j) {
                c += (u128)v[j] * b.v[i] + t[j];
                t[j] = (u64)c; c >>= 64;
            }
            c += t[N]; t[N] = (u64)c; t[N + 1] = (u64)(c >> 64);
            u64 m = t[0] * INV;
            c = (u128)m * MOD[0] + t[0]; c >>= 64;
            for (int j = 1; j < N; ++j) {
                c += (u128)m * MOD[j] + t[j];
                t[j - 1] = (u64)c; c >>= 64;
            }
            c += t[N]; t[N - 1] = (u64)c;
            t[N] = t[N + 1] + (u64)(c >> 64);
        }
        Field r;
        if (t[N] || cmp_limbs<N>(t, MOD) >= 0) sub_limbs<N>(r.v, t, MOD);
        else memcpy(r.v, t, sizeof(r.v));
        return r;
    }
    Field sqr() const { return (*this) * (*this); }
    Field &operator+=(const Field &b) { *this = *this + b; return *this; }
    Field &operator-=(const Field &b) { *this = *this - b; return *this; }
    Field &operator*=(const Field &b) { *this = *this * b; return *this; }

    // exponent: little-endian limbs
    Field pow(const u64 *e, int n) const {
        Field r = one();
        bool started = false;
        for (int i = n * 64 - 1; i >= 0; --i) {
            if (started) r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) { r = started ? r * (*this) : (*this); started = true; }
        }
        return r;
    }
    Field pow_u64(u64 e) const { return pow(&e, 1); }
    Field inv() const { return pow(PM2, N); }   // 0 -> 0
};
template <int N, int Tag> u64 Field<N, Tag>::MOD[N];
template <int N, int Tag> u64 Field<N, Tag>::INV;
template <int N, int Tag> Field<N, Tag> Field<N, Tag>::R1;
template <int N, int Tag> Field<N, Tag> Field<N, Tag>::R2;
template <int N, int Tag> u64 Field<N, Tag>::PM2[N];
template <int N, int Tag> u64 Field<N, Tag>::HALF[N];

typedef Field<6, 0> Fp;
typedef Field<4, 1> Fr;

static const u64 FP_MOD[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                              0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 FR_MOD[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL,
                              0x73eda753299d7d48ULL};
static const u64 BLS_X_ABS = 0xd201000000010000ULL;   // curve parameter x = -BLS_X_ABS

// gnark fr.BatchInvert semantics: zero entries stay zero (Montgomery's trick over the rest).
static inline void batch_invert(Fr *a, size_t n) {
    std::vector<Fr> pre(n);
    Fr acc = Fr::one();
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!a[i].is_zero()) acc = acc * a[i];
    }
    acc = acc.inv();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        Fr t = acc * pre[i];
        acc = acc * a[i];
        a[i] = t;
    }
}

// ---------------------------------------------------------------------------------------
// Fp2 = Fp[u]/(u^2+1)
// ---------------------------------------------------------------------------------------
struct Fp2 {
    Fp c0, c1;
    static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fp2 &b) const { return c0 == b.c0 && c1 == b.c1; }
    bool operator!=(const Fp2 &b) const { return !(*this == b); }
    Fp2 operator+(const Fp2 &b) const { return {c0 + b.c0, c1 + b.c1}; }
    Fp2 operator-(const Fp2 &b) const { return {c0 - b.c0, c1 - b.c1}; }
    Fp2 neg() const { return {c0.neg(), c1.neg()}; }
    Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fp2 conj() const { return {c0, c1.neg()}; }
    Fp2 operator*(const Fp2 &b) const {
        Fp a = c0 * b.c0, d = c1 * b.c1;
        Fp e = (c0 + c1) * (b.c0 + b.c1);
        return {a - d, e - a - d};
    }
    Fp2 mul_fp(const Fp &s) const { return {c0 * s, c1 * s}; }
    Fp2 sqr() const { return (*this) * (*this); }
    Fp2 mul_xi() const { return {c0 - c1, c0 + c1}; }   // * (1+u)
    Fp2 inv() const {
        Fp n = (c0.sqr() + c1.sqr()).inv();
        return {c0 * n, (c1 * n).neg()};
    }
    Fp2 pow(const u64 *e, int n) const {
        Fp2 r = one();
        for (int i = n * 64 - 1; i >= 0; --i) {
            r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * (*this);
        }
        return r;
    }
    // ZCash ordering: compare c1 first, then c0
    bool lex_largest() const { return c1.is_zero() ? c0.lex_largest() : c1.lex_largest(); }
    Fp2 &operator+=(const Fp2 &b) { *this = *this + b; return *this; }
    Fp2 &operator-=(const Fp2 &b) { *this = *this - b; return *this; }
    Fp2 &operator*=(const Fp2 &b) { *this = *this * b; return *this; }
};

// sqrt in Fp (p = 3 mod 4): candidate a^((p+1)/4); caller checks the square.
static inline bool fp_sqrt(Fp &out, const Fp &a) {
    static u64 e[6]; static bool init = false;
    if (!init) {   // (p+1)/4
        u64 one[6] = {1}, t[6];
        add_limbs<6>(t, Fp::MOD, one);
        for (int i = 0; i < 6; ++i) e[i] = (t[i] >> 2) | (i + 1 < 6 ? t[i + 1] << 62 : 0);
        init = true;
    }
    Fp s = a.pow(e, 6);
    if (s.sqr() != a) return false;
    out = s;
    return true;
}
// sqrt in Fp2 by the norm ("complex") method; result verified by squaring.
static inline bool fp2_sqrt(Fp2 &out, const Fp2 &a) {
    if (a.is_zero()) { out = a; return true; }
    Fp inv2 = Fp::from_u64(2).inv();
    if (a.c1.is_zero()) {
        Fp s;
        if (fp_sqrt(s, a.c0)) { out = {s, Fp::zero()}; return true; }
        if (fp_sqrt(s, a.c0.neg())) { out = {Fp::zero(), s}; return true; }
        return false;
    }
    Fp n = a.c0.sqr() + a.c1.sqr(), s;
    if (!fp_sqrt(s, n)) return false;
    Fp d = (a.c0 + s) * inv2, x0;
    if (!fp_sqrt(x0, d)) {
        d = (a.c0 - s) * inv2;
        if (!fp_sqrt(x0, d)) return false;
    }
    Fp x1 = a.c1 * (x0.dbl()).inv();
    Fp2 r = {x0, x1};
    if (r.sqr() != a) return false;
    out = r;
    return true;
}

// ---------------------------------------------------------------------------------------
// Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v),  xi = 1+u
// ---------------------------------------------------------------------------------------
struct Fp6 {
    Fp2 c0, c1, c2;
    static Fp6 zero() { return {Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
    static Fp6 one() { return {Fp2::one(), Fp2::zero(), Fp2::zero()}; }
    bool operator==(const Fp6 &b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
    Fp6 operator+(const Fp6 &b) const { return {c0 + b.c0, c1 + b.c1, c2 + b.c2}; }
    Fp6 operator-(const Fp6 &b) const { return {c0 - b.c0, c1 - b.c1, c2 - b.c2}; }
    Fp6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fp6 operator*(const Fp6 &b) const {
        Fp2 t0 = c0 * b.c0, t1 = c1 * b.c1, t2 = c2 * b.c2;
        Fp2 r0 = ((c1 + c2) * (b.c1 + b.c2) - t1 - t2).mul_xi() + t0;
        Fp2 r1 = (c0 + c1) * (b.c0 + b.c1) - t0 - t1 + t2.mul_xi();
        Fp2 r2 = (c0 + c2) * (b.c0 + b.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    Fp6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
    Fp6 inv() const {
        Fp2 t0 = c0.sqr() - (c1 * c2).mul_xi();
        Fp2 t1 = c2.sqr().mul_xi() - c0 * c1;
        Fp2 t2 = c1.sqr() - c0 * c2;
        Fp2 d = (c0 * t0 + (c2 * t1).mul_xi() + (c1 * t2).mul_xi()).inv();
        return {t0 * d, t1 * d, t2 * d};
    }
};

struct Fp12 {
    Fp6 c0, c1;
    static Fp12 one() { return {Fp6::one(), Fp6::zero()}; }
    bool operator==(const Fp12 &b) const { return c0 == b.c0 && c1 == b.c1; }
    Fp12 operator*(const Fp12 &b) const {
        Fp6 t0 = c0 * b.c0, t1 = c1 * b.c1;
        return {t0 + t1.mul_v(), (c0 + c1) * (b.c0 + b.c1) - t0 - t1};
    }
    Fp12 sqr() const { return (*this) * (*this); }
    Fp12 conj() const { return {c0, c1.neg()}; }
    Fp12 inv() const {
        Fp6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
    Fp12 pow_u64(u64 e) const {
        Fp12 r = one();
        for (int i = 63; i >= 0; --i) { r = r.sqr(); if ((e >> i) & 1) r = r * (*this); }
        return r;
    }
};

// Frobenius on Fp12: coefficients gamma[k][i] = xi^(i*(p^k-1)/6), computed at init.
struct Frob {
    Fp2 g1[6];   // xi^(i(p-1)/6)
};
static Frob &frob_tab() { static Frob f; return f; }
static inline void init_frob() {
    // (p-1)/6
    u64 e[6], one[6] = {1};
    sub_limbs<6>(e, Fp::MOD, one);
    u128 rem = 0;
    for (int i = 5; i >= 0; --i) { u128 cur = (rem << 64) | e[i]; e[i] = (u64)(cur / 6); rem = cur % 6; }
    Fp2 xi = {Fp::one(), Fp::one()};
    Fp2 g = xi.pow(e, 6);
    Frob &f = frob_tab();
    f.g1[0] = Fp2::one();
    for (int i = 1; i < 6; ++i) f.g1[i] = f.g1[i - 1] * g;
}
// x -> x^p.  Element = sum_{i,j} a_{ij} v^i w^j ; basis power of w is (2i + j).
static inline Fp12 frobenius(const Fp12 &a) {
    const Frob &f = frob_tab();
    Fp12 r;
    r.c0.c0 = a.c0.c0.conj();
    r.c0.c1 = a.c0.c1.conj() * f.g1[2];
    r.c0.c2 = a.c0.c2.conj() * f.g1[4];
    r.c1.c0 = a.c1.c0.conj() * f.g1[1];
    r.c1.c1 = a.c1.c1.conj() * f.g1[3];
    r.c1.c2 = a.c1.c2.conj() * f.g1[5];
    return r;
}

// ---------------------------------------------------------------------------------------
// Short Weierstrass y^2 = x^3 + b (a = 0), Jacobian coordinates, generic over the base field
// ---------------------------------------------------------------------------------------
template <class F> struct Aff {
    F x, y; bool inf;
    static Aff infinity() { return {F::zero(), F::zero(), true}; }
    Aff neg() const { return {x, y.neg(), inf}; }
    bool operator==(const Aff &b) const { return inf == b.inf && (inf || (x == b.x && y == b.y)); }
};
template <class F> struct Jac {
    F X, Y, Z;
    static Jac infinity() { return {F::one(), F::one(), F::zero()}; }
    static Jac from_affine(const Aff<F> &a) { return a.inf ? infinity() : Jac{a.x, a.y, F::one()}; }
    bool is_inf() const { return Z.is_zero(); }
    Jac neg() const { return {X, Y.neg(), Z}; }
    Jac dbl() const {
        if (is_inf()) return *this;
        F A = X.sqr(), B = Y.sqr(), C = B.sqr();
        F D = ((X + B).sqr() - A - C).dbl();
        F E = A.dbl() + A, Fq = E.sqr();
        F X3 = Fq - D.dbl();
        F Y3 = E * (D - X3) - C.dbl().dbl().dbl();
        F Z3 = (Y * Z).dbl();
        return {X3, Y3, Z3};
    }
    Jac add(const Jac &b) const {
        if (is_inf()) return b;
        if (b.is_inf()) return *this;
        F Z1Z1 = Z.sqr(), Z2Z2 = b.Z.sqr();
        F U1 = X * Z2Z2, U2 = b.X * Z1Z1;
        F S1 = Y * b.Z * Z2Z2, S2 = b.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return infinity();
        }
        F H = U2 - U1, I = H.dbl().sqr(), J = H * I, rr = (S2 - S1).dbl(), V = U1 * I;
        F X3 = rr.sqr() - J - V.dbl();
        F Y3 = rr * (V - X3) - (S1 * J).dbl();
        F Z3 = ((Z + b.Z).sqr() - Z1Z1 - Z2Z2) * H;
        return {X3, Y3, Z3};
    }
    Jac add_affine(const Aff<F> &b) const {
        if (b.inf) return *this;
        if (is_inf()) return from_affine(b);
        F Z1Z1 = Z.sqr();
        F U2 = b.x * Z1Z1, S2 = b.y * Z * Z1Z1;
        if (X == U2) {
            if (Y == S2) return dbl();
            return infinity();
        }
        F H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I, rr = (S2 - Y).dbl(), V = X * I;
        F X3 = rr.sqr() - J - V.dbl();
        F Y3 = rr * (V - X3) - (Y * J).dbl();
        F Z3 = (Z + H).sqr() - Z1Z1 - HH;
        return {X3, Y3, Z3};
    }
    Aff<F> to_affine() const {
        if (is_inf()) return Aff<F>::infinity();
        F zi = Z.inv(), zi2 = zi.sqr();
        return {X * zi2, Y * zi2 * zi, false};
    }
    // scalar: little-endian limbs (plain integer)
    Jac mul(const u64 *k, int n) const {
        // fixed 4-bit windows
        Jac tab[16];
        tab[0] = infinity();
        tab[1] = *this;
        for (int i = 2; i < 16; ++i) tab[i] = (i & 1) ? tab[i - 1].add(*this) : tab[i / 2].dbl();
        Jac r = infinity();
        for (int i = n * 16 - 1; i >= 0; --i) {
            r = r.dbl().dbl().dbl().dbl();
            unsigned d = (k[i / 16] >> ((i % 16) * 4)) & 15;
            if (d) r = r.add(tab[d]);
        }
        return r;
    }
};
typedef Aff<Fp> G1Affine;
typedef Jac<Fp> G1Jac;
typedef Aff<Fp2> G2Affine;
typedef Jac<Fp2> G2Jac;

static inline G1Jac g1_mul_fr(const G1Jac &p, const Fr &s) { u64 k[4]; s.to_limbs(k); return p.mul(k, 4); }
static inline G2Jac g2_mul_fr(const G2Jac &p, const Fr &s) { u64 k[4]; s.to_limbs(k); return p.mul(k, 4); }

static inline G1Affine g1_generator() {
    static const u64 gx[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                              0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
    static const u64 gy[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                              0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};
    return {Fp::from_limbs(gx), Fp::from_limbs(gy), false};
}

static inline bool g1_on_curve(const G1Affine &a) {
    if (a.inf) return true;
    return a.y.sqr() == a.x.sqr() * a.x + Fp::from_u64(4);
}
// definitional subgroup check: [r]P == O
static inline bool g1_in_subgroup(const G1Affine &a) {
    if (a.inf) return true;
    return G1Jac::from_affine(a).mul(FR_MOD, 4).is_inf();
}

// ---------------------------------------------------------------------------------------
// ZCash / gnark compressed encoding  (gnark-crypto ecc/bls12-381/marshal.go semantics:
// top three bits = compressed | infinity | y-is-lexicographically-largest)
// ---------------------------------------------------------------------------------------
enum DecodeStatus { DEC_OK = 0, DEC_BAD_ENCODING = 3, DEC_NOT_ON_CURVE = 4, DEC_NOT_IN_SUBGROUP = 5 };

static inline int g1_decompress(G1Affine &out, const uint8_t *b, bool subgroup_check) {
    unsigned m = b[0] >> 5;
    // 0b111, 0b011, 0b001 are invalid masks; 0b000 / 0b010 are uncompressed forms that need
    // 96 bytes (a 48-byte buffer is "short"): every one of them is an encoding error here.
    if (m != 4 && m != 5 && m != 6) return DEC_BAD_ENCODING;
    if (m == 6) {
        if (b[0] & 0x1f) return DEC_BAD_ENCODING;
        for (int i = 1; i < 48; ++i) if (b[i]) return DEC_BAD_ENCODING;
        out = G1Affine::infinity();
        return DEC_OK;
    }
    uint8_t xb[48]; memcpy(xb, b, 48); xb[0] &= 0x1f;
    Fp x;
    if (!Fp::from_bytes_be(x, xb)) return DEC_BAD_ENCODING;
    Fp y2 = x.sqr() * x + Fp::from_u64(4), y;
    if (!fp_sqrt(y, y2)) return DEC_NOT_ON_CURVE;
    if (y.lex_largest() != (m == 5)) y = y.neg();
    out = {x, y, false};
    if (subgroup_check && !g1_in_subgroup(out)) return DEC_NOT_IN_SUBGROUP;
    return DEC_OK;
}
static inline void g1_compress(uint8_t *b, const G1Affine &a) {
    if (a.inf) { memset(b, 0, 48); b[0] = 0xc0; return; }
    a.x.to_bytes_be(b);
    b[0] |= 0x80;
    if (a.y.lex_largest()) b[0] |= 0x20;
}
// G2: 96 bytes = x.c1 (with flags) || x.c0 ; no subgroup check (trusted_setup.go:130-133 parses
// with NoSubgroupChecks)
static inline int g2_decompress(G2Affine &out, const uint8_t *b) {
    unsigned m = b[0] >> 5;
    if (m != 4 && m != 5 && m != 6) return DEC_BAD_ENCODING;
    if (m == 6) { out = G2Affine::infinity(); return DEC_OK; }
    uint8_t xb[48]; memcpy(xb, b, 48); xb[0] &= 0x1f;
    Fp2 x;
    if (!Fp::from_bytes_be(x.c1, xb)) return DEC_BAD_ENCODING;
    if (!Fp::from_bytes_be(x.c0, b + 48)) return DEC_BAD_ENCODING;
    Fp2 b2 = {Fp::from_u64(4), Fp::from_u64(4)};
    Fp2 y2 = x.sqr() * x + b2, y;
    if (!fp2_sqrt(y, y2)) return DEC_NOT_ON_CURVE;
    if (y.lex_largest() != (m == 5)) y = y.neg();
    out = {x, y, false};
    return DEC_OK;
}

// ---------------------------------------------------------------------------------------
// optimal-ate pairing product check:  prod e(P_i, Q_i) == 1
// Miller loop on the M-type twist with affine line functions
//   l(P) * w^3 = (lambda*xT - yT) + (-lambda*xP) v + (yP) v w
// ---------------------------------------------------------------------------------------
static inline Fp12 line_eval(const Fp2 &lambda, const G2Affine &T, const G1Affine &P) {
    Fp12 l;
    l.c0 = {lambda * T.x - T.y, lambda.mul_fp(P.x).neg(), Fp2::zero()};
    l.c1 = {Fp2::zero(), Fp2{P.y, Fp::zero()}, Fp2::zero()};
    return l;
}
static inline Fp12 miller_loop(const G1Affine &P, const G2Affine &Q) {
    Fp12 f = Fp12::one();
    if (P.inf || Q.inf) return f;
    G2Affine T = Q;
    for (int i = 62; i >= 0; --i) {   // BLS_X_ABS has its top bit at position 63
        f = f.sqr();
        Fp2 lambda = (T.x.sqr().dbl() + T.x.sqr()) * T.y.dbl().inv();
        f = f * line_eval(lambda, T, P);
        Fp2 x3 = lambda.sqr() - T.x.dbl();
        Fp2 y3 = lambda * (T.x - x3) - T.y;
        T = {x3, y3, false};
        if ((BLS_X_ABS >> i) & 1) {
            Fp2 lam2 = (Q.y - T.y) * (Q.x - T.x).inv();
            f = f * line_eval(lam2, T, P);
            Fp2 x4 = lam2.sqr() - T.x - Q.x;
            Fp2 y4 = lam2 * (T.x - x4) - T.y;
            T = {x4, y4, false};
        }
    }
    return f.conj();   // x < 0
}
static inline Fp12 final_exp(const Fp12 &f0) {
    // easy part: f^((p^6-1)(p^2+1))
    Fp12 f = f0.conj() * f0.inv();
    f = frobenius(frobenius(f)) * f;
    // hard part, exponent 3*(p^4-p^2+1)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3
    auto powx = [](const Fp12 &a) { return a.pow_u64(BLS_X_ABS).conj(); };   // a^x, x<0 (cyclotomic: inverse = conj)
    Fp12 a = powx(f) * f.conj();          // f^(x-1)
    a = powx(a) * a.conj();               // f^((x-1)^2)
    Fp12 b = powx(a) * frobenius(a);      // ^(x+p)
    Fp12 c = powx(powx(b)) * frobenius(frobenius(b)) * b.conj();   // ^(x^2+p^2-1)
    return c * f.sqr() * f;
}
static inline bool pairing_check(const G1Affine *P, const G2Affine *Q, int n) {
    Fp12 f = Fp12::one();
    for (int i = 0; i < n; ++i) f = f * miller_loop(P[i], Q[i]);
    return final_exp(f) == Fp12::one();
}

static inline void init_all() {
    static bool done = false;
    if (done) return;
    Fp::init(FP_MOD);
    Fr::init(FR_MOD);
    init_frob();
    done = true;
}

}  // namespace ko
