// Modular inversion by an optimised binary GCD (after T. Pornin, "Optimized Binary GCD for Modular
// Inversion", ePrint 2020/972): 31 plain binary-GCD steps at a time are run on 64-bit
// APPROXIMATIONS of (a, b) (their top 33 and low 31 bits) while the 2x2 update matrix is tracked in
// small integers; the matrix is then applied to the full-size numbers with one multi-limb
// multiply-add per operand.  ~25 outer rounds for a 381-bit modulus, a few hundred instructions
// each, against 380 squarings + ~90 products (~190 k instructions) for the Fermat chain it
// replaces (gnark-crypto's fp.Element.Inverse is also a binary-GCD variant; call sites in the
// reference: fr.BatchInvert at internal/domain/domain.go:95,213, internal/kzg/kzg_prove.go:95,137).
//
// Most of the work is shifts, adds and selects, i.e. it runs on the ALU pipe next to the IMAD-bound
// curve arithmetic.  The code is plain C++ on 32-bit limbs, compiled for host AND device: the host
// build is unit-tested against Python's pow(x, -1, m) on the CPU (tests/test_bingcd.py).
//
// Contract: m odd, 0 < y < m, gcd(y, m) = 1 (m is prime at every call site).  Plain integers in and
// out, little-endian 32-bit limbs; y = 0 returns 0.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define KZG_HD __host__ __device__ __forceinline__
#else
#define KZG_HD inline
#endif

namespace kzg {

template <int N> struct BinGcd {
    // r = (s0 * |f0| * x0  +  s1 * |f1| * x1), s = +-1, as a two's-complement (N+2)-limb number
    static KZG_HD void lin2(uint32_t *r, const uint32_t *x0, int64_t f0, const uint32_t *x1, int64_t f1) {
        const uint64_t a0 = (uint64_t)(f0 < 0 ? -f0 : f0), a1 = (uint64_t)(f1 < 0 ? -f1 : f1);   // <= 2^31
        uint32_t p[N + 2], q[N + 2];
        uint64_t c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { c += (uint64_t)x0[i] * a0; p[i] = (uint32_t)c; c >>= 32; }
        p[N] = (uint32_t)c; p[N + 1] = 0;
        c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { c += (uint64_t)x1[i] * a1; q[i] = (uint32_t)c; c >>= 32; }
        q[N] = (uint32_t)c; q[N + 1] = 0;
        // r = (+-p) + (+-q): conditional negation by xor/carry, all in (N+2)-limb two's complement
        const uint32_t m0 = f0 < 0 ? 0xffffffffu : 0u, m1 = f1 < 0 ? 0xffffffffu : 0u;
        uint64_t cp = m0 & 1u, cq = m1 & 1u, cs = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N + 2; ++i) {
            cp += (uint64_t)(p[i] ^ m0); uint32_t pi = (uint32_t)cp; cp >>= 32;
            cq += (uint64_t)(q[i] ^ m1); uint32_t qi = (uint32_t)cq; cq >>= 32;
            cs += (uint64_t)pi + qi; r[i] = (uint32_t)cs; cs >>= 32;
        }
    }
    // x = |r| >> 31 (exact), returns true if r was negative
    static KZG_HD bool abs_shift31(uint32_t *x, const uint32_t *r) {
        const bool neg = (r[N + 1] >> 31) != 0;
        const uint32_t msk = neg ? 0xffffffffu : 0u;
        uint32_t t[N + 2];
        uint64_t c = msk & 1u;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N + 2; ++i) { c += (uint64_t)(r[i] ^ msk); t[i] = (uint32_t)c; c >>= 32; }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) x[i] = (t[i] >> 31) | (t[i + 1] << 1);
        return neg;
    }
    // u = (u * f0 + v * f1) / 2^31 mod m for u, v in [0, m); negative coefficients act on m - u, m - v
    static KZG_HD void lin2_mod(uint32_t *out, const uint32_t *u, int64_t f0, const uint32_t *v, int64_t f1, const uint32_t *m, uint32_t m_ninv31) {
        uint32_t un[N], vn[N];
        neg_mod(un, u, m); neg_mod(vn, v, m);
        const uint64_t a0 = (uint64_t)(f0 < 0 ? -f0 : f0), a1 = (uint64_t)(f1 < 0 ? -f1 : f1);
        uint32_t x0[N], x1[N];                                          // element-wise selects keep everything in registers
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { x0[i] = f0 < 0 ? un[i] : u[i]; x1[i] = f1 < 0 ? vn[i] : v[i]; }
        uint32_t t[N + 2];
        uint64_t c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) {                                   // t = x0 a0 + x1 a1  (< 2 m 2^31)
            uint64_t lo = (uint64_t)x0[i] * a0, hi = (uint64_t)x1[i] * a1;
            c += (lo & 0xffffffffu) + (hi & 0xffffffffu);
            t[i] = (uint32_t)c;
            c = (c >> 32) + (lo >> 32) + (hi >> 32);
        }
        t[N] = (uint32_t)c; t[N + 1] = (uint32_t)(c >> 32);
        // make t divisible by 2^31: add k m with k = t * (-1/m) mod 2^31
        const uint64_t k = (uint64_t)((t[0] * m_ninv31) & 0x7fffffffu);
        c = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { c += (uint64_t)m[i] * k + t[i]; t[i] = (uint32_t)c; c >>= 32; }
        c += t[N]; t[N] = (uint32_t)c; c >>= 32;
        t[N + 1] += (uint32_t)c;
        uint32_t w[N + 1];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N + 1; ++i) w[i] = (t[i] >> 31) | (t[i + 1] << 1);        // < 3 m
        // two conditional subtractions of m
        for (int rep = 0; rep < 2; ++rep) {
            uint32_t d[N + 1];
            uint64_t br = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int i = 0; i < N + 1; ++i) {
                uint64_t s = (uint64_t)w[i] - (i < N ? m[i] : 0u) - br;
                d[i] = (uint32_t)s; br = (s >> 32) & 1u;
            }
            if (!br) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
                for (int i = 0; i < N + 1; ++i) w[i] = d[i];
            }
        }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) out[i] = w[i];
    }
    static KZG_HD void neg_mod(uint32_t *r, const uint32_t *x, const uint32_t *m) {     // m - x, and 0 -> 0
        uint32_t nz = 0;
        uint64_t br = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { nz |= x[i]; uint64_t s = (uint64_t)m[i] - x[i] - br; r[i] = (uint32_t)s; br = (s >> 32) & 1u; }
        if (!nz) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int i = 0; i < N; ++i) r[i] = 0;
        }
    }
    static KZG_HD int bitlen(const uint32_t *x) {
        int n = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) if (x[i]) {
#ifdef __CUDA_ARCH__
            n = 32 * i + 32 - __clz((int)x[i]);
#else
            n = 32 * i + 32 - __builtin_clz(x[i]);
#endif
        }
        return n;
    }
    // bits [pos, pos + 64) of x (pos >= 0), zero-extended
    static KZG_HD uint64_t bits64(const uint32_t *x, int pos) {
        const int w = pos >> 5, s = pos & 31;
        uint32_t l0 = 0, l1 = 0, l2 = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { if (i == w) l0 = x[i]; if (i == w + 1) l1 = x[i]; if (i == w + 2) l2 = x[i]; }
        uint64_t lo = ((uint64_t)l1 << 32) | l0;
        return s ? (lo >> s) | ((uint64_t)l2 << (64 - s)) : lo;
    }

    // out = y^-1 mod m.  m_ninv31 = -m^-1 mod 2^31.  Returns the number of outer rounds used.
    static KZG_HD int inverse(uint32_t *out, const uint32_t *y, const uint32_t *m, uint32_t m_ninv31) {
        uint32_t a[N], b[N], u[N], v[N];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) { a[i] = y[i]; b[i] = m[i]; u[i] = 0; v[i] = 0; }
        u[0] = 1;
        int rounds = 0;
        for (; rounds < 2 * (32 * N + 30) / 31 + 2; ++rounds) {
            uint32_t nz = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int i = 0; i < N; ++i) nz |= a[i];
            if (!nz) break;
            // approximations: low 31 bits and top 33 bits of a and b at a common alignment
            const int la = bitlen(a), lb = bitlen(b), n = la > lb ? la : lb;
            uint64_t xa, xb;
            if (n <= 64) { xa = bits64(a, 0); xb = bits64(b, 0); }
            else {
                xa = (uint64_t)(a[0] & 0x7fffffffu) | ((bits64(a, n - 33) & 0x1ffffffffULL) << 31);
                xb = (uint64_t)(b[0] & 0x7fffffffu) | ((bits64(b, n - 33) & 0x1ffffffffULL) << 31);
            }
            int64_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
            for (int i = 0; i < 31; ++i) {
                if (xa & 1u) {
                    if (xa < xb) {
                        uint64_t t = xa; xa = xb; xb = t;
                        int64_t s = f0; f0 = f1; f1 = s;
                        s = g0; g0 = g1; g1 = s;
                    }
                    xa -= xb; f0 -= f1; g0 -= g1;
                }
                xa >>= 1; f1 <<= 1; g1 <<= 1;
            }
            // (a, b) <- (a f0 + b g0, a f1 + b g1) / 2^31, made non-negative
            uint32_t ra[N + 2], rb[N + 2];
            lin2(ra, a, f0, b, g0);
            lin2(rb, a, f1, b, g1);
            if (abs_shift31(a, ra)) { f0 = -f0; g0 = -g0; }
            if (abs_shift31(b, rb)) { f1 = -f1; g1 = -g1; }
            // (u, v) <- the same combination modulo m (with the division by 2^31 done modulo m)
            uint32_t nu[N], nv[N];
            lin2_mod(nu, u, f0, v, g0, m, m_ninv31);
            lin2_mod(nv, u, f1, v, g1, m, m_ninv31);
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
            for (int i = 0; i < N; ++i) { u[i] = nu[i]; v[i] = nv[i]; }
        }
        // a == 0, b == gcd == 1: v y == 1 (mod m)
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < N; ++i) out[i] = v[i];
        return rounds;
    }
};

}  // namespace kzg
