// Montgomery arithmetic over N x 32-bit limbs for sm_100a, carry chains in registers.
//
// Replaces (on device) what the reference gets from gnark-crypto's fp.Element / fr.Element
// (go.mod:8; call sites e.g. internal/domain/fft.go:109-144, internal/kzg/kzg_prove.go:81-111).
//
// Multiplication is the operand-scanning "even/odd" scheme: the running sum T is kept as
// T = even + 2^32 * odd, so every 32x32->64 product (mad.lo.cc / madc.hi.cc pair) lands on a
// 64-bit aligned limb pair of one of the two accumulators and each row is two independent carry
// chains on the IMAD pipe.  After each Montgomery step the low limb is zero and the two
// accumulators swap roles instead of shifting.  IMAD count per product: N*(4N+1)
// (588 for Fp, 264 for Fr) -- the figure SURVEY.md section 8(d) uses for the roofline.
//
// Credit: the even/odd-accumulator row scheme and the names of its helpers (mul_n, cmad_n, madc_n_rshift,
// mad_n_redc, final_sub) are those of supranational/sppark's mont_t.cuh (Apache-2.0), restated here from its
// published design; the dedicated squaring (sqr) and reduce_wide are this repository's own.
//
// All values are kept fully reduced in [0, p).  p must leave >= 2 spare bits in the top limb
// (true for both BLS12-381 moduli) so that 2p fits in N limbs.
#pragma once
#include <cstdint>

namespace kzg {

// ---- PTX carry-chain primitives (volatile: program order of the chain is preserved) ---------
__device__ __forceinline__ uint32_t ptx_add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t ptx_mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t ptx_madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t ptx_madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t ptx_madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

// Field parameters live in __constant__ memory; P::mod() etc. return pointers into it.
template <class P> struct Mont {
    static constexpr int N = P::N;
    uint32_t v[N];

    // ---- chains ---------------------------------------------------------------------------
    // acc[j], acc[j+1] = a[j] * b   for even j  (a points at the limb that pairs with acc[0])
    static __device__ __forceinline__ void mul_n(uint32_t *acc, const uint32_t *a, uint32_t b) {
#pragma unroll
        for (int j = 0; j < N; j += 2) { acc[j] = ptx_mul_lo(a[j], b); acc[j + 1] = ptx_mul_hi(a[j], b); }
    }
    // acc += sum_{j even} a[j]*b << 32j ; one carry chain, carry-out left in CC
    static __device__ __forceinline__ void cmad_n(uint32_t *acc, const uint32_t *a, uint32_t b) {
        acc[0] = ptx_mad_lo_cc(a[0], b, acc[0]);
        acc[1] = ptx_madc_hi_cc(a[0], b, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) { acc[j] = ptx_madc_lo_cc(a[j], b, acc[j]); acc[j + 1] = ptx_madc_hi_cc(a[j], b, acc[j + 1]); }
    }
    // acc = (acc >> 64) + sum_{j even} a[j]*b << 32j, carry-in from CC
    static __device__ __forceinline__ void madc_n_rshift(uint32_t *acc, const uint32_t *a, uint32_t b) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) { acc[j] = ptx_madc_lo_cc(a[j], b, acc[j + 2]); acc[j + 1] = ptx_madc_hi_cc(a[j], b, acc[j + 3]); }
        acc[N - 2] = ptx_madc_lo_cc(a[N - 2], b, 0);
        acc[N - 1] = ptx_madc_hi(a[N - 2], b, 0);
    }
    // one row: (even + 2^32 odd) = ((even + 2^32 odd) + a*bi + m*p) / 2^32 with roles swapped by caller
    static __device__ __forceinline__ void mad_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi, bool first) {
        const uint32_t *MOD = P::mod();
        if (first) {
            mul_n(odd, a + 1, bi);
            mul_n(even, a, bi);
        } else {
            even[0] = ptx_add_cc(even[0], odd[1]);
            madc_n_rshift(odd, a + 1, bi);
            cmad_n(even, a, bi);
            odd[N - 1] = ptx_addc(odd[N - 1], 0);
        }
        uint32_t mi = even[0] * P::inv();
        cmad_n(odd, MOD + 1, mi);
        cmad_n(even, MOD, mi);
        odd[N - 1] = ptx_addc(odd[N - 1], 0);
    }

    // r = r - p if r >= p   (r < 2p on entry)
    static __device__ __forceinline__ void final_sub(uint32_t *r) {
        const uint32_t *MOD = P::mod();
        uint32_t t[N];
        t[0] = ptx_sub_cc(r[0], MOD[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) t[i] = ptx_subc_cc(r[i], MOD[i]);
        uint32_t borrow = ptx_subc(0, 0);   // 0xffffffff if r < p
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = borrow ? r[i] : t[i];
    }

    // Montgomery reduction of a 2N-limb value (< p * 2^(32N)) with reduction-only rows
    static __device__ __forceinline__ Mont reduce_wide(const uint32_t *T) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; ++i) even[i] = T[i];
#pragma unroll
        for (int r = 0; r < N; r += 2) {
            redc_row(even, odd, r == 0);
            redc_row(odd, even, false);
        }
        even[0] = ptx_add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) even[i] = ptx_addc_cc(even[i], odd[i + 1]);
        even[N - 1] = ptx_addc(even[N - 1], 0);
        even[0] = ptx_add_cc(even[0], T[N]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) even[i] = ptx_addc_cc(even[i], T[N + i]);
        even[N - 1] = ptx_addc(even[N - 1], T[2 * N - 1]);
        final_sub(even);
        Mont r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = even[i];
        return r;
    }

    // ---- field ops ------------------------------------------------------------------------
    static __device__ __forceinline__ Mont mul(const Mont &a, const Mont &b) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_n_redc(even, odd, a.v, b.v[i], i == 0);
            mad_n_redc(odd, even, a.v, b.v[i + 1], false);
        }
        // result = even + (odd >> 32)
        even[0] = ptx_add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) even[i] = ptx_addc_cc(even[i], odd[i + 1]);
        even[N - 1] = ptx_addc(even[N - 1], 0);
        final_sub(even);
        Mont r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = even[i];
        return r;
    }
    // (a*b + c*d) / R mod p with ONE Montgomery reduction: every row adds a*b_i, c*d_i and m*p, i.e. 3N wide products
    // per row pair instead of the 4N of two separate products (432 + 12 instead of 576 + 24 IMADs for Fp).  ONLY for moduli
    // with 3p < 2^(32N) (Fp: p ~ 0.10 * 2^384; NOT Fr, whose 255-bit r leaves one spare bit): the running sum then stays
    // below 3p*2^32 < 2^(32N+32), so no limb chain overflows, and the result is below (2p^2 + Rp)/R < 1.21 p: one
    // conditional subtraction, as in mul().
    static __device__ __forceinline__ void mad2_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi, const uint32_t *c, uint32_t di, bool first) {
        const uint32_t *MOD = P::mod();
        if (first) {
            mul_n(odd, a + 1, bi);
            mul_n(even, a, bi);
        } else {
            even[0] = ptx_add_cc(even[0], odd[1]);
            madc_n_rshift(odd, a + 1, bi);
            cmad_n(even, a, bi);
            odd[N - 1] = ptx_addc(odd[N - 1], 0);
        }
        cmad_n(odd, c + 1, di);
        cmad_n(even, c, di);
        odd[N - 1] = ptx_addc(odd[N - 1], 0);
        uint32_t mi = even[0] * P::inv();
        cmad_n(odd, MOD + 1, mi);
        cmad_n(even, MOD, mi);
        odd[N - 1] = ptx_addc(odd[N - 1], 0);
    }
    static __device__ __forceinline__ Mont mul_add_mul(const Mont &a, const Mont &b, const Mont &c, const Mont &d) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad2_n_redc(even, odd, a.v, b.v[i], c.v, d.v[i], i == 0);
            mad2_n_redc(odd, even, a.v, b.v[i + 1], c.v, d.v[i + 1], false);
        }
        even[0] = ptx_add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) even[i] = ptx_addc_cc(even[i], odd[i + 1]);
        even[N - 1] = ptx_addc(even[N - 1], 0);
        final_sub(even);
        Mont r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = even[i];
        return r;
    }
    // Dedicated squaring: 2N^2+N -> N(N+1)/2 + N^2 + N wide products (234 vs 300 for N = 12).
    //  1. T = a^2 as 2N limbs.  Cross products a_i*a_j (i<j) land on limbs (i+j, i+j+1); they are
    //     accumulated into two arrays by the parity of i+j so every row is two contiguous carry
    //     chains (row i: j = i+1, i+3, ... and j = i+2, i+4, ...).  The limb after a chain's end has
    //     not been touched by earlier rows, so one trailing addc absorbs the chain's carry.
    //  2. T = 2*(E+O) + sum a_i^2 2^(64 i).
    //  3. Montgomery-reduce the low half with "reduction-only" rows (the multiplication rows of
    //     mul() without the a*b_i term), add the high half, one conditional subtraction.
    static __device__ __forceinline__ Mont sqr(const Mont &a) {
        uint32_t E[2 * N], O[2 * N];
#pragma unroll
        for (int i = 0; i < 2 * N; ++i) { E[i] = 0; O[i] = 0; }
#pragma unroll
        for (int i = 0; i < N - 1; ++i) {
            // chain over j = i+1, i+3, ... : positions i+j odd-offset from 2i -> parity of (2i+1) = odd -> array O
            {
                int j = i + 1, s = i + j;
                O[s] = ptx_mad_lo_cc(a.v[i], a.v[j], O[s]);
                O[s + 1] = ptx_madc_hi_cc(a.v[i], a.v[j], O[s + 1]);
#pragma unroll
                for (j = i + 3; j < N; j += 2) {
                    s = i + j;
                    O[s] = ptx_madc_lo_cc(a.v[i], a.v[j], O[s]);
                    O[s + 1] = ptx_madc_hi_cc(a.v[i], a.v[j], O[s + 1]);
                }
                // last j of this class
                int jl = i + 1 + 2 * ((N - 1 - (i + 1)) / 2);
                int t = i + jl + 1;
                if (t + 1 < 2 * N) O[t + 1] = ptx_addc(O[t + 1], 0); else (void)ptx_addc(0, 0);
            }
            // chain over j = i+2, i+4, ... : even positions -> array E
            if (i + 2 < N) {
                int j = i + 2, s = i + j;
                E[s] = ptx_mad_lo_cc(a.v[i], a.v[j], E[s]);
                E[s + 1] = ptx_madc_hi_cc(a.v[i], a.v[j], E[s + 1]);
#pragma unroll
                for (j = i + 4; j < N; j += 2) {
                    s = i + j;
                    E[s] = ptx_madc_lo_cc(a.v[i], a.v[j], E[s]);
                    E[s + 1] = ptx_madc_hi_cc(a.v[i], a.v[j], E[s + 1]);
                }
                int jl = i + 2 + 2 * ((N - 1 - (i + 2)) / 2);
                int t = i + jl + 1;
                if (t + 1 < 2 * N) E[t + 1] = ptx_addc(E[t + 1], 0); else (void)ptx_addc(0, 0);
            }
        }
        // T = E + O
        uint32_t T[2 * N];
        T[0] = ptx_add_cc(E[0], O[0]);
#pragma unroll
        for (int i = 1; i < 2 * N - 1; ++i) T[i] = ptx_addc_cc(E[i], O[i]);
        T[2 * N - 1] = ptx_addc(E[2 * N - 1], O[2 * N - 1]);
        // T = 2T (cross terms count twice)
        T[0] = ptx_add_cc(T[0], T[0]);
#pragma unroll
        for (int i = 1; i < 2 * N - 1; ++i) T[i] = ptx_addc_cc(T[i], T[i]);
        T[2 * N - 1] = ptx_addc(T[2 * N - 1], T[2 * N - 1]);
        // T += sum a_i^2 << 64 i   (one contiguous chain)
        T[0] = ptx_mad_lo_cc(a.v[0], a.v[0], T[0]);
        T[1] = ptx_madc_hi_cc(a.v[0], a.v[0], T[1]);
#pragma unroll
        for (int i = 1; i < N; ++i) {
            T[2 * i] = ptx_madc_lo_cc(a.v[i], a.v[i], T[2 * i]);
            T[2 * i + 1] = ptx_madc_hi_cc(a.v[i], a.v[i], T[2 * i + 1]);
        }
        return reduce_wide(T);
    }
    // one reduction-only row: (even + 2^32 odd) <- ((even + 2^32 odd) + m*p) / 2^32, roles swapped by the caller
    static __device__ __forceinline__ void redc_row(uint32_t *even, uint32_t *odd, bool first) {
        const uint32_t *MOD = P::mod();
        if (first) {
            uint32_t mi = even[0] * P::inv();
            mul_n(odd, MOD + 1, mi);
            cmad_n(even, MOD, mi);
            odd[N - 1] = ptx_addc(odd[N - 1], 0);
        } else {
            even[0] = ptx_add_cc(even[0], odd[1]);
            uint32_t mi = even[0] * P::inv();
            madc_n_rshift(odd, MOD + 1, mi);
            cmad_n(even, MOD, mi);
            odd[N - 1] = ptx_addc(odd[N - 1], 0);
        }
    }

    static __device__ __forceinline__ Mont add(const Mont &a, const Mont &b) {
        Mont r;
        r.v[0] = ptx_add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.v[i] = ptx_addc_cc(a.v[i], b.v[i]);
        r.v[N - 1] = ptx_addc(a.v[N - 1], b.v[N - 1]);
        final_sub(r.v);
        return r;
    }
    static __device__ __forceinline__ Mont sub(const Mont &a, const Mont &b) {
        const uint32_t *MOD = P::mod();
        Mont r;
        r.v[0] = ptx_sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) r.v[i] = ptx_subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = ptx_subc(0, 0);   // all ones if a < b
        r.v[0] = ptx_add_cc(r.v[0], MOD[0] & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) r.v[i] = ptx_addc_cc(r.v[i], MOD[i] & borrow);
        r.v[N - 1] = ptx_addc(r.v[N - 1], MOD[N - 1] & borrow);
        return r;
    }
    static __device__ __forceinline__ Mont neg(const Mont &a) { return sub(zero(), a); }
    static __device__ __forceinline__ Mont dbl(const Mont &a) { return add(a, a); }

    static __device__ __forceinline__ Mont zero() {
        Mont r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = 0;
        return r;
    }
    static __device__ __forceinline__ Mont one() {   // R mod p
        Mont r; const uint32_t *o = P::r1();
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = o[i];
        return r;
    }
    __device__ __forceinline__ bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    static __device__ __forceinline__ bool eq(const Mont &a, const Mont &b) {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= a.v[i] ^ b.v[i];
        return o == 0;
    }
    // plain integer (< p) -> Montgomery form, and back
    static __device__ __forceinline__ Mont to_mont(const Mont &a) {
        Mont r2; const uint32_t *o = P::r2();
#pragma unroll
        for (int i = 0; i < N; ++i) r2.v[i] = o[i];
        return mul(a, r2);
    }
    static __device__ __forceinline__ Mont from_mont(const Mont &a) {
        Mont o = zero(); o.v[0] = 1;
        return mul(a, o);
    }
    // plain integer comparison a >= b (both plain limbs)
    static __device__ __forceinline__ bool geq_limbs(const uint32_t *a, const uint32_t *b) {
        ptx_sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) ptx_subc_cc(a[i], b[i]);
        return ptx_subc(0, 0) == 0;
    }
};

}  // namespace kzg
