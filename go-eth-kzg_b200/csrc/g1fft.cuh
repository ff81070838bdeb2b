// Batched size-128 G1 FFTs for FK20, one kernel launch per radix-2 stage over the whole batch.
//
// Replaces the two group FFTs of internal/kzg_multi/fk20/fk20.go:76-124 (-> toeplitz.go:113-125,
// internal/domain/fft.go:39-83: IFFT of the 128 MSM sums, keep 64, zero-pad, FFT).
//
// Work decomposition: thread (butterfly b, blob n) of stage s.  All threads of a block share the
// butterfly, so the twiddle w^t -- a compile-time constant of the protocol -- is block-uniform and
// its scalar multiplication is a fixed op list in constant memory ("double k times, add +-T[i]"),
// with no per-lane digit scanning and no divergence.  Blocks whose twiddle is 1 only do the two
// additions; they are numbered last so they fill the tail of the launch.  The working set
// (128 Jacobian points per blob, 18 KB) lives in global memory and is L2-resident for a chunk
// of 1024 blobs; every stage reads and writes it once.
//
// [w^t]P (jac_mul_prog): GLV split w^t = k1 + k2*lambda, both halves in width-5 NAF (~42 additions
// and ~128 doublings).  The table of odd multiples P, 3P, .. 15P is built on the curve isomorphic
// to E in which 2P is affine (mixed additions, 8M+3S), rescaled to one common Z and then used as
// an AFFINE table on that isomorphic curve (a = 0 formulas do not involve b), so the main loop
// also runs on mixed additions.  The result is mapped back by Z *= Z_common * Z_2P.
#pragma once
#include "g1.cuh"

namespace kzg {

#ifndef KZG_G1FFT_TPB
#define KZG_G1FFT_TPB 128   // measured on B200 (1024 blobs, split 8): 32 -> 37.8 ms, 64 -> 34.7 ms, 128 -> 33.1 ms
#endif
__device__ __constant__ uint16_t TW_PROG[128][KZG_TW_PROG_LEN];   // uploaded from H_TW_PROG (constants.inc)
__device__ __constant__ uint16_t DENSE_PROG[65][KZG_TW_PROG_LEN]; // uploaded from H_DENSE_PROG: scalars of the dense small-batch transform

// 16-byte vector moves of field elements / points (all records are 16-byte aligned)
__device__ __forceinline__ Fp ld_fp(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = q[0], b = q[1], c = q[2];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
    return r;
}
__device__ __forceinline__ void st_fp(Fp *p, const Fp &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
    q[2] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
}
__device__ __forceinline__ G1J ld_jac(const G1J *p) { G1J r; r.X = ld_fp(&p->X); r.Y = ld_fp(&p->Y); r.Z = ld_fp(&p->Z); return r; }
__device__ __forceinline__ void st_jac(G1J *p, const G1J &r) { st_fp(&p->X, r.X); st_fp(&p->Y, r.Y); st_fp(&p->Z, r.Z); }

// a += (x2, y2), Jacobian + affine, 8M + 3S.  a must not be the point at infinity.
template <class M_ = MulCall> __device__ __forceinline__ void jac_madd(G1J &a, const Fp &x2, const Fp &y2) {
    Fp Z1Z1 = M_::sqr(a.Z);
    Fp U2 = M_::mul(x2, Z1Z1), S2 = M_::mul(M_::mul(y2, a.Z), Z1Z1);
    Fp H = Fp::sub(U2, a.X), r = Fp::sub(S2, a.Y);
    if (H.is_zero()) {
        if (r.is_zero()) { a.X = x2; a.Y = y2; a.Z = Fp::one(); jac_dbl<M_>(a); }
        else a.Z = Fp::zero();
        return;
    }
    Fp HH = M_::sqr(H), HHH = M_::mul(H, HH), V = M_::mul(a.X, HH);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(r), HHH), Fp::dbl(V));
    a.Y = M_::msub(r, Fp::sub(V, X3), a.Y, HHH);
    a.X = X3;
    a.Z = M_::mul(a.Z, H);
}

// a += b, both Jacobian (add-2007-bl shape, 11M + 5S), every degenerate case handled
template <class M_ = MulCall> __device__ __forceinline__ void jac_add_full(G1J &a, const G1J &b) {
    if (b.Z.is_zero()) return;
    if (a.Z.is_zero()) { a = b; return; }
    Fp Z1Z1 = M_::sqr(a.Z), Z2Z2 = M_::sqr(b.Z);
    Fp U1 = M_::mul(a.X, Z2Z2), U2 = M_::mul(b.X, Z1Z1);
    Fp S1 = M_::mul(M_::mul(a.Y, b.Z), Z2Z2), S2 = M_::mul(M_::mul(b.Y, a.Z), Z1Z1);
    Fp H = Fp::sub(U2, U1), r = Fp::sub(S2, S1);
    if (H.is_zero()) {
        if (r.is_zero()) jac_dbl<M_>(a);
        else a.Z = Fp::zero();
        return;
    }
    Fp HH = M_::sqr(H), HHH = M_::mul(H, HH), V = M_::mul(U1, HH);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(r), HHH), Fp::dbl(V));
    a.Y = M_::msub(r, Fp::sub(V, X3), S1, HHH);
    a.X = X3;
    a.Z = M_::mul(M_::mul(a.Z, b.Z), H);
}
static __device__ __noinline__ void jac_add_ool(G1J *a, const G1J *b) { jac_add_full<MulCallLazy>(*a, *b); }

// *pp = [w_128^t] *pp, t block-uniform in 1..127.  Infinity in -> infinity out (Z = 0 propagates
// through Z_2P).  Scripts/check_tw_prog.py is the big-integer model of this routine.
// M_ = MulCallLazy (Y3 = r (V - X3) - Y1 HHH under one Montgomery reduction, mont.cuh: mul_add_mul)
template <class M_> static __device__ __noinline__ void jac_mul_prog_at(G1J *pp, const uint16_t *prog);
template <class M_ = MulCallLazy> static __device__ __forceinline__ void jac_mul_prog(G1J *pp, int t) { jac_mul_prog_at<M_>(pp, TW_PROG[t]); }
// *pp = [k] *pp for the fixed scalar whose op list (constant memory, block-uniform) is prog
template <class M_> static __device__ __noinline__ void jac_mul_prog_at(G1J *pp, const uint16_t *prog) {
    Fp tx[8], ty[8], tbx[8], zr[8];
    Fp ZC;                                     // Z_common * Z_2P
    {
        G1J P = *pp;
        G1J D = P;
        jac_dbl<M_>(D);
        Fp C2 = M_::sqr(D.Z), C3 = M_::mul(C2, D.Z);
        G1J T; T.X = M_::mul(P.X, C2); T.Y = M_::mul(P.Y, C3); T.Z = P.Z;      // P on the curve where 2P = (D.X, D.Y) is affine
        tx[0] = T.X; ty[0] = T.Y;
#pragma unroll 1
        for (int k = 1; k < 8; ++k) {          // T_k = T_{k-1} + 2P, raw mixed addition; zr[k] = Z_k / Z_{k-1}
            Fp Z1Z1 = M_::sqr(T.Z);
            Fp U2 = M_::mul(D.X, Z1Z1), S2 = M_::mul(M_::mul(D.Y, T.Z), Z1Z1);
            Fp H = Fp::sub(U2, T.X), r = Fp::sub(S2, T.Y);
            Fp HH = M_::sqr(H), HHH = M_::mul(H, HH), V = M_::mul(T.X, HH);
            Fp X3 = Fp::sub(Fp::sub(M_::sqr(r), HHH), Fp::dbl(V));
            T.Y = M_::msub(r, Fp::sub(V, X3), T.Y, HHH);
            T.X = X3;
            T.Z = M_::mul(T.Z, H);
            tx[k] = T.X; ty[k] = T.Y; zr[k] = H;
        }
        ZC = M_::mul(T.Z, D.Z);
    }
    {
        Fp beta;
#pragma unroll
        for (int i = 0; i < 12; ++i) beta.v[i] = FP_BETA[i];
        tbx[7] = M_::mul(tx[7], beta);
        Fp s = zr[7];                          // s = Z_7 / Z_k
#pragma unroll 1
        for (int k = 6; k >= 0; --k) {
            Fp s2 = M_::sqr(s);
            Fp x = M_::mul(tx[k], s2);
            tx[k] = x;
            tbx[k] = M_::mul(x, beta);
            ty[k] = M_::mul(ty[k], M_::mul(s2, s));
            if (k) s = M_::mul(s, zr[k]);
        }
    }
    const int n_ops = prog[0] & 255, trailing = prog[0] >> 8;
    G1J acc;
    {
        uint32_t op = prog[1];
        int idx = (op >> 8) & 7;
        acc.X = (op & 0x1000u) ? tbx[idx] : tx[idx];
        acc.Y = (op & 0x0800u) ? Fp::neg(ty[idx]) : ty[idx];
        acc.Z = Fp::one();
    }
#pragma unroll 1
    for (int k = 2; k <= n_ops; ++k) {
        uint32_t op = prog[k];
#pragma unroll 1
        for (int d = op & 255; d > 0; --d) jac_dbl<M_>(acc);
        int idx = (op >> 8) & 7;
        Fp ex = (op & 0x1000u) ? tbx[idx] : tx[idx];
        Fp ey = ty[idx];
        if (op & 0x0800u) ey = Fp::neg(ey);
        jac_madd<M_>(acc, ex, ey);
    }
#pragma unroll 1
    for (int d = trailing; d > 0; --d) jac_dbl<M_>(acc);
    acc.Z = M_::mul(acc.Z, ZC);
    *pp = acc;
}

// ---- the same twiddle multiplication with every pair of independent field products in ONE out-of-line body (g1.cuh: fp_sqr2_ni,
// fp_mul2_ni, fp_sqrmul_ni, fp_mammul_ni).  The arithmetic is identical (same products, same reductions: bit-exact); what changes is the
// depth of the dependent chain a warp walks -- 4 calls per doubling instead of 7, 5 per mixed addition instead of 11 -- so that each warp
// keeps about twice as many carry chains in flight.  A stage of the transform over 1024 blobs is only ~10 warps per SM of work
// (DESIGN.md section 8), too few single-chain warps to saturate the multiply pipe; compiled for 2 CTAs per SM (255 registers: the dual
// bodies take up to 72 argument registers) it is still ~8 dual-chain warps.
// MEASURED (profiles/r02_rejected_dual_and_lane_split.md): 34.5 ms against 33.8 ms for the single-chain kernel at 3 CTAs per SM -- the SASS of the
// dual bodies is interleaved as intended, but a warp gains only ~11 %, which does not pay for the third CTA.  Kept as the run-time tunable
// "g1fft_dual" (default off); outputs are byte-identical.
__device__ __forceinline__ void jac_dbl2(G1J &p) {            // dbl-2009-l, a = 0; no-op on infinity (Z = 0 stays 0)
    FpPair q = fp_sqr2_ni(p.X, p.Y);
    const Fp A = q.a, B = q.b;
    q = fp_sqr2_ni(B, Fp::add(p.X, B));
    const Fp C = q.a;
    const Fp D = Fp::dbl(Fp::sub(Fp::sub(q.b, A), C));
    const Fp E = Fp::add(Fp::dbl(A), A);
    q = fp_sqrmul_ni(E, p.Y, p.Z);
    p.X = Fp::sub(q.a, Fp::dbl(D));
    p.Z = Fp::dbl(q.b);
    p.Y = Fp::sub(fp_mul_ni(E, Fp::sub(D, p.X)), Fp::dbl(Fp::dbl(Fp::dbl(C))));
}
// a += (x2, y2), Jacobian + affine; a must not be the point at infinity.  hout (optional): H = Z3 / Z1
__device__ __forceinline__ void jac_madd2(G1J &a, const Fp &x2, const Fp &y2, Fp *hout = nullptr) {
    FpPair q = fp_sqrmul_ni(a.Z, y2, a.Z);
    const Fp Z1Z1 = q.a;
    q = fp_mul2_ni(x2, Z1Z1, q.b, Z1Z1);
    const Fp H = Fp::sub(q.a, a.X), r = Fp::sub(q.b, a.Y);
    if (hout) *hout = H;
    if (H.is_zero()) {
        if (r.is_zero()) { a.X = x2; a.Y = y2; a.Z = Fp::one(); jac_dbl2(a); }
        else a.Z = Fp::zero();
        return;
    }
    q = fp_sqr2_ni(H, r);
    const Fp HH = q.a, r2 = q.b;
    q = fp_mul2_ni(H, HH, a.X, HH);
    const Fp HHH = q.a, V = q.b;
    const Fp X3 = Fp::sub(Fp::sub(r2, HHH), Fp::dbl(V));
    q = fp_mammul_ni(r, Fp::sub(V, X3), Fp::neg(a.Y), HHH, a.Z, H);
    a.X = X3; a.Y = q.a; a.Z = q.b;
}
static __device__ __noinline__ void jac_mul_prog_dual_at(G1J *pp, const uint16_t *prog) {
    Fp tx[8], ty[8], tbx[8], zr[8];
    Fp ZC;                                     // Z_common * Z_2P
    {
        G1J P = *pp;
        G1J D = P;
        jac_dbl2(D);
        const Fp C2 = fp_sqr_ni(D.Z);
        FpPair q = fp_mul2_ni(C2, D.Z, P.X, C2);
        G1J T; T.X = q.b; T.Y = fp_mul_ni(P.Y, q.a); T.Z = P.Z;      // P on the curve where 2P = (D.X, D.Y) is affine
        tx[0] = T.X; ty[0] = T.Y;
#pragma unroll 1
        for (int k = 1; k < 8; ++k) {          // T_k = T_{k-1} + 2P, raw mixed addition (never degenerate: k P != +-2P); zr[k] = Z_k / Z_{k-1}
            FpPair u = fp_sqrmul_ni(T.Z, D.Y, T.Z);
            const Fp Z1Z1 = u.a;
            u = fp_mul2_ni(D.X, Z1Z1, u.b, Z1Z1);
            const Fp H = Fp::sub(u.a, T.X), r = Fp::sub(u.b, T.Y);
            u = fp_sqr2_ni(H, r);
            const Fp HH = u.a, r2 = u.b;
            u = fp_mul2_ni(H, HH, T.X, HH);
            const Fp HHH = u.a, V = u.b;
            const Fp X3 = Fp::sub(Fp::sub(r2, HHH), Fp::dbl(V));
            u = fp_mammul_ni(r, Fp::sub(V, X3), Fp::neg(T.Y), HHH, T.Z, H);
            T.X = X3; T.Y = u.a; T.Z = u.b;
            tx[k] = T.X; ty[k] = T.Y; zr[k] = H;
        }
        ZC = fp_mul_ni(T.Z, D.Z);
    }
    {
        Fp beta;
#pragma unroll
        for (int i = 0; i < 12; ++i) beta.v[i] = FP_BETA[i];
        tbx[7] = fp_mul_ni(tx[7], beta);
        Fp s = zr[7];                          // s = Z_7 / Z_k
#pragma unroll 1
        for (int k = 6; k >= 0; --k) {
            FpPair u = fp_sqrmul_ni(s, s, zr[k ? k : 1]);          // s^2 and the next s (unused after k = 0)
            const Fp s2 = u.a, s_next = u.b;
            u = fp_mul2_ni(tx[k], s2, s2, s);
            const Fp x = u.a;
            tx[k] = x;
            u = fp_mul2_ni(x, beta, ty[k], u.b);
            tbx[k] = u.a; ty[k] = u.b;
            s = s_next;
        }
    }
    const int n_ops = prog[0] & 255, trailing = prog[0] >> 8;
    G1J acc;
    {
        uint32_t op = prog[1];
        int idx = (op >> 8) & 7;
        acc.X = (op & 0x1000u) ? tbx[idx] : tx[idx];
        acc.Y = (op & 0x0800u) ? Fp::neg(ty[idx]) : ty[idx];
        acc.Z = Fp::one();
    }
#pragma unroll 1
    for (int k = 2; k <= n_ops; ++k) {
        uint32_t op = prog[k];
#pragma unroll 1
        for (int d = op & 255; d > 0; --d) jac_dbl2(acc);
        int idx = (op >> 8) & 7;
        Fp ex = (op & 0x1000u) ? tbx[idx] : tx[idx];
        Fp ey = ty[idx];
        if (op & 0x0800u) ey = Fp::neg(ey);
        jac_madd2(acc, ex, ey);
    }
#pragma unroll 1
    for (int d = trailing; d > 0; --d) jac_dbl2(acc);
    acc.Z = fp_mul_ni(acc.Z, ZC);
    *pp = acc;
}

// butterfly number of the y-th block of a stage: non-trivial twiddles first
__device__ __forceinline__ int g1fft_butterfly(int y, int log_half) {
    const int half = 1 << log_half;
    if (half == 1) return y;
    const int heavy = 64 - (64 >> log_half);
    if (y < heavy) { int g = y / (half - 1), j = y - g * (half - 1) + 1; return (g << log_half) + j; }
    return (y - heavy) << log_half;
}

// One radix-2 stage over all blobs.  grid = (ceil(blobs / TPB), 64), block = TPB.
//   DIT (bit-reversed in, natural out):  y *= w; (x, y) <- (x + y, x - y)
//   DIF (natural in, bit-reversed out):  (x, y) <- (x + y, (x - y) * w)
// SRC_XYZZ: the stage reads the MSM sums (XYZZ) instead of the working set.  DST_XYZZ: it writes
// XYZZ points for k_finalize_g1.  ONLY_SUM: x - y is not needed (inverse transform, last stage,
// upper half discarded: toeplitz.go:124).  UPPER_ZERO: y is the zero padding (fk20.go:82-85).
template <bool DIT, bool INVERSE, bool SRC_XYZZ, bool DST_XYZZ, bool ONLY_SUM, bool UPPER_ZERO, bool DUAL>
static __global__ void __launch_bounds__(KZG_G1FFT_TPB, DUAL ? 2 : 3) k_g1fft_stage(const G1 *__restrict__ src_xyzz, G1J *__restrict__ work, G1 *__restrict__ dst_xyzz,
                                                                      const int32_t *__restrict__ status, int nblobs, int log_half) {
    const int blob = blockIdx.x * KZG_G1FFT_TPB + threadIdx.x;
    if (blob >= nblobs || status[blob] != ST_OK) return;
    const int bf = g1fft_butterfly(blockIdx.y, log_half);
    const int half = 1 << log_half, j = bf & (half - 1);
    const int i0 = ((bf >> log_half) << (log_half + 1)) + j, i1 = i0 + half;
    int t = j * (64 >> log_half);
    if (INVERSE) t = (128 - t) & 127;
    G1J *w0 = work + (size_t)blob * 128 + i0, *w1 = work + (size_t)blob * 128 + i1;
    G1J x, y;
    if (!UPPER_ZERO) {
        if (SRC_XYZZ) { G1 q = src_xyzz[(size_t)blob * 128 + i1]; y = jac_from_xyzz(q); }
        else y = ld_jac(w1);
    }
    if (DIT) {
        if (t) { if (DUAL) jac_mul_prog_dual_at(&y, TW_PROG[t]); else jac_mul_prog<MulCallLazy>(&y, t); }
        if (SRC_XYZZ) { G1 q = src_xyzz[(size_t)blob * 128 + i0]; x = jac_from_xyzz(q); }
        else x = ld_jac(w0);
        G1J s = x;
        jac_add_ool(&s, &y);
        if (!ONLY_SUM) { y.Y = Fp::neg(y.Y); jac_add_ool(&x, &y); }
        y = x; x = s;
    } else {
        if (SRC_XYZZ) { G1 q = src_xyzz[(size_t)blob * 128 + i0]; x = jac_from_xyzz(q); }
        else x = ld_jac(w0);
        if (UPPER_ZERO) y = x;
        else {
            G1J s = x;
            jac_add_ool(&s, &y);
            y.Y = Fp::neg(y.Y); jac_add_ool(&x, &y);
            y = x; x = s;
        }
        if (t) { if (DUAL) jac_mul_prog_dual_at(&y, TW_PROG[t]); else jac_mul_prog<MulCallLazy>(&y, t); }
    }
    if (DST_XYZZ) {
        dst_xyzz[(size_t)blob * 128 + i0] = jac_to_xyzz(x);
        dst_xyzz[(size_t)blob * 128 + i1] = jac_to_xyzz(y);
    } else {
        st_jac(w0, x);
        if (!ONLY_SUM) st_jac(w1, y);
    }
}

// ---- dense form for SMALL batches -----------------------------------------------------------------------
// The staged transform above is a chain of 14 launches, each one ~1.1 ms of dependent doublings per thread: 16 ms however
// small the batch (the reference is called with ONE blob, api_eip7594.go:28).  IFFT_128 -> keep 64 -> zero-pad -> FFT_128
// is the circulant map
//     P[m] = 64 S[m] + sum_{q : m - q odd} c_(m-q) S[q],      c_d = sum_{k<64} w^(dk) = 2 / (1 - w^d)
// with 65 fixed scalars, so for a few blobs all 128 x 65 products [c]S[q] are computed AT ONCE (one scalar
// multiplication deep instead of 13; 13x the arithmetic of the staged form, which an otherwise idle GPU has to spare)
// and each output is a 65-term sum.  S arrives bit-reversed (position j holds frequency brp7(j)), P leaves bit-reversed.
//   k_g1dense_mul: grid (ceil(blobs * 128 / 32), 65), block 32: thread = (blob, position), blockIdx.y = scalar slot
//   k_g1dense_sum: grid (128, blobs), block 64: one output per block, tree over its 64 + 1 terms
static __global__ void __launch_bounds__(32) k_g1dense_mul(const G1 *__restrict__ sums, G1J *__restrict__ prod, const int32_t *__restrict__ status, int nblobs) {
    const int idx = blockIdx.x * 32 + threadIdx.x, blob = idx >> 7, pos = idx & 127, slot = blockIdx.y;
    if (blob >= nblobs || status[blob] != ST_OK) return;
    G1 s = sums[(size_t)blob * 128 + pos];
    G1J y = jac_from_xyzz(s);
    jac_mul_prog_at<MulCallLazy>(&y, DENSE_PROG[slot]);
    const int q = (int)(__brev((unsigned)pos) >> 25);
    st_jac(prod + ((size_t)blob * 65 + slot) * 128 + q, y);
}
static __global__ void __launch_bounds__(64) k_g1dense_sum(const G1J *__restrict__ prod, G1 *__restrict__ dst_xyzz, const int32_t *__restrict__ status) {
    __shared__ G1J sm[64];
    const int blob = blockIdx.y, t = threadIdx.x;
    if (status[blob] != ST_OK) return;
    const int m = (int)(__brev((unsigned)blockIdx.x) >> 25);
    const G1J *base = prod + (size_t)blob * 65 * 128;
    const int q = 2 * t + (1 - (m & 1)), d = (m - q) & 127;
    G1J acc = ld_jac(base + (size_t)(d >> 1) * 128 + q);
    if (t == 0) { G1J e = ld_jac(base + (size_t)64 * 128 + m); jac_add_ool(&acc, &e); }
    sm[t] = acc;
    __syncthreads();
    for (int s = 32; s > 0; s >>= 1) {
        if (t < s) jac_add_ool(&sm[t], &sm[t + s]);
        __syncthreads();
    }
    if (t == 0) dst_xyzz[(size_t)blob * 128 + blockIdx.x] = jac_to_xyzz(sm[0]);
}

// ---- two-level form for the batches just above the dense form's range (25 .. 32 blobs per GPU) --------------------------------
// Between the dense form (13 x the arithmetic, one multiplication deep: pays up to ~24 blobs) and the staged form (the minimum
// arithmetic, 14 multiplications deep: 16 ms however small the batch) sits the Cooley-Tukey split 128 = 16 x 8 with every level
// done densely: all products of a level at once, then short sums.  With w = w_128, S[f] the MSM sums (frequency f at position
// brp7(f)), h = IFFT_128(S)[0..63] (1/128 folded into the scalars), P = FFT_128(h, zero-padded), P[k] leaving at position brp7(k):
//   I1:  A[f2][m1] = sum_{f1 < 16} S[8 f1 + f2] w^(-8 f1 m1)          128 outputs x 16 terms
//   I2:  h[m]      = sum_{f2 < 8}  A[f2][m mod 16] w^(-f2 m), m < 64    64 outputs x  8 terms
//   F1:  B[b][k1]  = sum_{a < 8}   h[8 a + b] w^(8 a k1)              128 outputs x  8 terms
//   F2:  P[k]      = sum_{b < 8}   B[b][k mod 16] w^(b k)             128 outputs x  8 terms
// ~4 100 non-trivial twiddle multiplications per blob (staged: 642, dense: 8 320), four multiplications deep.  MEASURED: 9.2 ms for
// 25-32 blobs (staged: 16.2 ms), 15.5 ms for 33-64 (a full wave of the GPU is 57 k twiddle multiplications = 2.5 ms of the multiply
// pipe, and 64 blobs x 4 100 are 4.6 waves), 22 ms for 96: it pays up to 32 blobs (kzg_lane::g1_two_level_max).  Same group elements
// as the other two forms, so the compressed proofs are byte-identical.
//   k_g1lvl_mul: grid (ceil(blobs / 32), n_out * R), block 32: thread = blob, block = (output o, term j): the twiddle is block-uniform
//   k_g1lvl_sum: one thread per (blob, output): R-term sum
__host__ __device__ __forceinline__ int g1_brp7(int x) {      // 7-bit reversal (host + device: the index maps below are unit-tested on the CPU)
    x = ((x & 0x55) << 1) | ((x >> 1) & 0x55); x = ((x & 0x33) << 2) | ((x >> 2) & 0x33); x = ((x & 0x0f) << 4) | ((x >> 4) & 0x0f);
    return (x >> 1) & 127;
}
__host__ __device__ __forceinline__ void g1lvl_term(int level, int o, int j, int &src, int &e) {
    switch (level) {
    case 0: { const int f2 = o >> 4, m1 = o & 15; src = g1_brp7(8 * j + f2); e = (128 - ((8 * j * m1) & 127)) & 127; break; }   // I1: src = position of S[8 j + f2]
    case 1: { src = (j << 4) + (o & 15); e = (128 - ((j * o) & 127)) & 127; break; }                                                                  // I2: A[j][o mod 16]
    case 2: { const int b = o >> 4, k1 = o & 15; src = 8 * j + b; e = (8 * j * k1) & 127; break; }                                                      // F1: h[8 j + b]
    default: { src = (j << 4) + (o & 15); e = (j * o) & 127; break; }                                                                                  // F2: B[j][o mod 16]
    }
}
// ---- the same idea one step further for 33 .. ~160 blobs: 128 = 4 x 4 x 4 x 2, eight levels of 4- or 2-term sums ------------------
// Splitting the size-16 transform of the form above as 4 x 4 and its size-8 transform (with the twiddle) as 4 x 2 (indices
// f = 8 (4 a + b) + f2, f2 = 2 p + q on the input side, k = k1 + 16 u + 64 v, k1 = c + 4 d on the output side, W = w^(-1) for the
// inverse and w for the forward transform):
//   B[f2][b][c]  = sum_{a < 4} x[8 (4 a + b) + f2] W^(32 a c)              (forward: a < 2, the upper half of h is the zero padding)
//   A[f2][k1]    = sum_{b < 4} B[f2][b][k1 mod 4]  W^(8 b k1)
//   C[k1][q][u]  = sum_{p < 4} A[2 p + q][k1]      W^(2 p (k1 + 16 u))
//   X[k]         = sum_{q < 2} C[k mod 16][q][(k / 16) mod 4] W^(q k)        (inverse: k < 64 only)
// ~2 000 non-trivial twiddle multiplications per blob (half of the 16 x 8 form, three times the staged form), eight multiplications deep
// (staged: fourteen).  Levels 0-3 are the inverse transform, 4-7 the forward one (outputs x terms: 128x4, 128x4, 128x4, 64x2, 128x2, 128x4, 128x4, 128x2).
// MEASURED: 12.6 ms for 33-64 blobs (staged 16.2, 16 x 8 form 15.5), 18.5 ms for 65-96 (a level is one partial wave whose blocks share the
// pipe: its duration is the longest twiddle multiplication under that contention, not the average), so it is used for 33-64 blobs.
__host__ __device__ __forceinline__ void g1lvl8_term(int level, int o, int j, int &src, int &e) {
    const int l = level & 3;
    int ex;
    if (l == 0) { const int f2 = o >> 4, b = (o >> 2) & 3, c = o & 3; src = 8 * (4 * j + b) + f2; ex = 32 * j * c; }
    else if (l == 1) { const int f2 = o >> 4, k1 = o & 15; src = f2 * 16 + j * 4 + (k1 & 3); ex = 8 * j * k1; }
    else if (l == 2) { const int k1 = o >> 3, q = (o >> 2) & 1, u = o & 3; src = (2 * j + q) * 16 + k1; ex = 2 * j * (k1 + 16 * u); }
    else { const int k1 = o & 15, u = (o >> 4) & 3; src = k1 * 8 + j * 4 + u; ex = j * o; }
    ex &= 127;
    e = level < 4 ? (128 - ex) & 127 : ex;
    if (level == 0) src = g1_brp7(src);      // the MSM sums arrive bit-reversed: frequency f at position brp7(f)
}
template <int LEVEL>
static __global__ void __launch_bounds__(32) k_g1lvl8_mul(const G1 *__restrict__ sums, const G1J *__restrict__ in, G1J *__restrict__ prod, const int32_t *__restrict__ status,
                                                      int nblobs, int R) {
    const int blob = blockIdx.x * 32 + threadIdx.x;
    if (blob >= nblobs || status[blob] != ST_OK) return;
    const int o = blockIdx.y / R, j = blockIdx.y - o * R;
    int src, e;
    g1lvl8_term(LEVEL, o, j, src, e);
    G1J y;
    if (LEVEL == 0) { G1 q = sums[(size_t)blob * 128 + src]; y = jac_from_xyzz(q); }
    else y = ld_jac(in + (size_t)blob * 128 + src);
    if (e) jac_mul_prog_at<MulCallLazy>(&y, TW_PROG[e]);
    st_jac(prod + ((size_t)blob * gridDim.y + blockIdx.y), y);
}

template <int LEVEL>
static __global__ void __launch_bounds__(32) k_g1lvl_mul(const G1 *__restrict__ sums, const G1J *__restrict__ in, G1J *__restrict__ prod, const int32_t *__restrict__ status,
                                                     int nblobs, int R) {
    const int blob = blockIdx.x * 32 + threadIdx.x;
    if (blob >= nblobs || status[blob] != ST_OK) return;
    const int o = blockIdx.y / R, j = blockIdx.y - o * R;
    int src, e;
    g1lvl_term(LEVEL, o, j, src, e);
    G1J y;
    if (LEVEL == 0) { G1 q = sums[(size_t)blob * 128 + src]; y = jac_from_xyzz(q); }
    else y = ld_jac(in + (size_t)blob * 128 + src);
    if (e) jac_mul_prog_at<MulCallLazy>(&y, TW_PROG[e]);
    st_jac(prod + ((size_t)blob * gridDim.y + blockIdx.y), y);
}
// out[blob][o] = sum_j prod[blob][o][j]; LAST: written as XYZZ at position brp7(o) for k_finalize_g1
template <bool LAST>
static __global__ void __launch_bounds__(64) k_g1lvl_sum(const G1J *__restrict__ prod, G1J *__restrict__ out, G1 *__restrict__ dst_xyzz, const int32_t *__restrict__ status,
                                                     int nblobs, int n_out, int R) {
    const int idx = blockIdx.x * 64 + threadIdx.x, blob = idx / n_out, o = idx - blob * n_out;
    if (blob >= nblobs || status[blob] != ST_OK) return;
    const G1J *p = prod + ((size_t)blob * n_out + o) * R;
    G1J acc = ld_jac(p);
#pragma unroll 1
    for (int j = 1; j < R; ++j) { G1J t = ld_jac(p + j); jac_add_ool(&acc, &t); }
    if (LAST) dst_xyzz[(size_t)blob * 128 + (int)(__brev((unsigned)o) >> 25)] = jac_to_xyzz(acc);
    else st_jac(out + (size_t)blob * 128 + o, acc);
}

}  // namespace kzg
