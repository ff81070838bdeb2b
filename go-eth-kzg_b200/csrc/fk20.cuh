// EIP-7594 cell + FK20 multi-proof pipeline on device.
//
// Replaces, for a batch of blobs at once:
//   api_eip7594.go:28-52          ComputeCellsAndKZGProofs
//   internal/kzg_multi/fk20/fk20.go:58-124   ComputeExtendedPolynomial / ComputeMultiOpenProof
//   internal/kzg_multi/fk20/toeplitz.go:95-125  BatchMulAggregation
//   internal/domain/fft.go:23-92  FftG1 / IfftG1 (one full scalar multiplication per twiddle)
//
// Order bookkeeping (why there is no bit-reversal kernel anywhere):
//   blob (brp order) --DIT inverse--> coefficients (natural)                         [k_blob_ifft]
//   coefficients * w_8192^j --DIF forward--> odd-indexed extension in brp order
//        == cells 64..127; cells 0..63 are the blob itself                           [k_coset_fft_cells]
//   64 circulant rows --DIF forward FFT-128--> scalars at brp positions q            [k_fk20_rows]
//   FK20 table stored with rows in the same brp order q (built by a DIF G1 FFT)      [k_g1_fft128_dif]
//   u'[q] = MSM_64(row q) --DIT inverse G1 FFT--> h natural; keep 64; pad            [k_g1fft_stage, g1fft.cuh]
//        --DIF forward G1 FFT--> proofs in brp order (= fk20.go:88-90 BitReverse)
// The 1/128 of the inverse G1 FFT is folded into the Fr scalars (the map is linear).
//
// The per-blob G1 FFTs of the proving path live in g1fft.cuh (one launch per stage over the batch).
// The shared-memory FFT below (per-lane GLV digits, signed base-16, 128 doublings + <= 64 additions
// per twiddle) is kept for context creation (FK20 table rows) and the cell verifier.
#pragma once
#include "ntt.cuh"
#include "constants.inc"

namespace kzg {

// ---- Fr side -----------------------------------------------------------------------------
// grid = blobs, block = NTT_THREADS, dynamic smem = 4096*32 B
#define KZG_NTT_THREADS 512
static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_blob_ifft(const uint8_t *__restrict__ blobs, Fr *__restrict__ coeffs,
                                                               int32_t *__restrict__ status, const Fr *__restrict__ roots, Fr inv_n) {
    extern __shared__ uint32_t sm[];
    const int blob = blockIdx.x, tid = threadIdx.x;
    const uint8_t *src = blobs + (size_t)blob * 131072;
    int bad = 0;
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.v[i] = FR_R2[i];
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x;
        load_be32(x.v, src + i * 32);
        if (!fr_is_canonical(x.v)) bad = 1;
        sm_store<4096>(sm, i, fr_mul_ni(x, r2));
    }
    bad = __syncthreads_or(bad);
    if (bad) { if (tid == 0) status[blob] = ST_NON_CANONICAL_SCALAR; return; }
    if (tid == 0) status[blob] = ST_OK;
    ntt_smem<12, true, true>(sm, roots, tid, KZG_NTT_THREADS);
    Fr *dst = coeffs + (size_t)blob * 4096;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) st_fr(dst + i, fr_mul_ni(sm_load<4096>(sm, i), inv_n));
}

// cells[blob] = blob bytes || brp(coset FFT) ; failed blobs get zeros
static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_coset_fft_cells(const Fr *__restrict__ coeffs, const uint8_t *__restrict__ blobs,
                                                                     uint8_t *__restrict__ cells, const int32_t *__restrict__ status,
                                                                     const Fr *__restrict__ roots) {
    extern __shared__ uint32_t sm[];
    const int blob = blockIdx.x, tid = threadIdx.x;
    uint4 *out = reinterpret_cast<uint4 *>(cells + (size_t)blob * 262144);
    if (status[blob] != ST_OK) {
        for (int i = tid; i < 262144 / 16; i += KZG_NTT_THREADS) out[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    const uint4 *in = reinterpret_cast<const uint4 *>(blobs + (size_t)blob * 131072);
    for (int i = tid; i < 131072 / 16; i += KZG_NTT_THREADS) out[i] = in[i];
    const Fr *src = coeffs + (size_t)blob * 4096;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x = ld_fr(src + i);
        if (i) x = fr_mul_ni(x, ld_fr(roots + i));   // coset shift w_8192^i
        sm_store<4096>(sm, i, x);
    }
    __syncthreads();
    ntt_smem<12, false, false>(sm, roots, tid, KZG_NTT_THREADS);
    Fr one_plain = Fr::zero(); one_plain.v[0] = 1;
    uint8_t *dst = cells + (size_t)blob * 262144 + 131072;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x = fr_mul_ni(sm_load<4096>(sm, i), one_plain);
        store_be32(dst + i * 32, x.v);
    }
}

// grid = (64 rows, blobs), block = 64.  scalars[blob][q][row] = plain( FFT128(circulant row)[brp q] / 128 )
static __global__ void __launch_bounds__(64) k_fk20_rows(const Fr *__restrict__ coeffs, uint32_t *__restrict__ scalars,
                                                  const int32_t *__restrict__ status, const Fr *__restrict__ roots, Fr inv128_plain) {
    __shared__ uint32_t sm[128 * 8];
    const int row = blockIdx.x, blob = blockIdx.y, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    const Fr *cf = coeffs + (size_t)blob * 4096;
    // toeplitzRows[row][m] = rev[row + 64 m] = coeffs[4095 - row - 64 m]   (fk20.go:106-110)
    // circulant = [r0, 0 x 64, r63, r62, ..., r1]                           (toeplitz.go:17-29)
    for (int i = tid; i < 128; i += 64) {
        Fr x = Fr::zero();
        if (i == 0) x = ld_fr(cf + 4095 - row);
        else if (i > 64) x = ld_fr(cf + 4095 - row - 64 * (128 - i));
        sm_store<128>(sm, i, x);
    }
    __syncthreads();
    ntt_smem<7, false, false>(sm, roots, tid, 64);
    uint32_t *dst = scalars + (size_t)blob * 8192 * 8;
    for (int q = tid; q < 128; q += 64) {
        Fr x = fr_mul_ni(sm_load<128>(sm, q), inv128_plain);   // Montgomery * plain = plain
        uint4 *o = reinterpret_cast<uint4 *>(dst + ((size_t)q * 64 + row) * 8);
        o[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
        o[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    }
}

// ---- G1 side -----------------------------------------------------------------------------
// P = [w_128^t] P with the precomputed GLV digits dig[2][KZG_GLV_DIGITS] (top window first)
template <class M_> static __device__ __noinline__ void g1_mul_twiddle_t(G1 *pp, const int8_t *__restrict__ dig) {
    G1 P = *pp;
    if (P.is_inf()) return;
    G1J base = jac_from_xyzz(P);
    G1JT tab[8];
    tab[0] = jac_cache(base);
    { G1J d = base; jac_dbl(d); tab[1] = jac_cache(d); }
#pragma unroll 1
    for (int i = 2; i < 8; ++i) {
        G1J t; t.X = tab[i - 1].X; t.Y = tab[i - 1].Y; t.Z = tab[i - 1].Z;
        jac_add(t, tab[0]);
        tab[i] = jac_cache(t);
    }
    Fp beta;
#pragma unroll
    for (int i = 0; i < 12; ++i) beta.v[i] = FP_BETA[i];
    G1J acc; acc.X = Fp::zero(); acc.Y = Fp::zero(); acc.Z = Fp::zero();
#pragma unroll 1
    for (int w = 0; w < KZG_GLV_DIGITS; ++w) {
        if (w) {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) jac_dbl<M_>(acc);
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            int d = dig[h * KZG_GLV_DIGITS + w];
            if (d == 0) continue;
            G1JT t = tab[(d < 0 ? -d : d) - 1];
            if (h) t.X = fp_mul_ni(t.X, beta);
            if (d < 0) t.Y = Fp::neg(t.Y);
            jac_add<M_>(acc, t);
        }
    }
    *pp = jac_to_xyzz(acc);
}

static __device__ __forceinline__ void g1_mul_twiddle(G1 *pp, const int8_t *__restrict__ dig) { g1_mul_twiddle_t<MulCall>(pp, dig); }
// radix-2 stages of a size-128 G1 FFT on points in shared memory, 64 threads.
//   DIT: bit-reversed in -> natural out.  DIF: natural in -> bit-reversed out.
template <bool DIT, bool INVERSE, class M_ = MulCall>
__device__ __forceinline__ void g1_fft128_smem(G1 *pts, const int8_t *__restrict__ digits, int tid) {
#pragma unroll 1
    for (int s = 0; s < 7; ++s) {
        const int log_half = DIT ? s : 6 - s;
        const int half = 1 << log_half;
        // butterfly index: in the half == 2 stage every other butterfly has a trivial twiddle; put all
        // non-trivial ones into warp 0 so warp 1 skips the multiplication instead of idling half its lanes
        int bf = tid;
        if (half == 2) bf = tid < 32 ? 2 * tid + 1 : 2 * (tid - 32);
        int j = bf & (half - 1);
        int i0 = ((bf >> log_half) << (log_half + 1)) + j, i1 = i0 + half;
        int t = j * (64 >> log_half);
        if (INVERSE) t = (128 - t) & 127;
        const int8_t *dg = digits + (size_t)t * 2 * KZG_GLV_DIGITS;
        G1 x = pts[i0], y = pts[i1];
        if (DIT) {
            if (t) g1_mul_twiddle_t<M_>(&y, dg);
            G1 s0 = x; g1_add(s0, y);
            y.neg_inplace(); g1_add(x, y);
            pts[i0] = s0; pts[i1] = x;
        } else {
            G1 s0 = x; g1_add(s0, y);
            y.neg_inplace(); g1_add(x, y);
            if (t) g1_mul_twiddle_t<M_>(&x, dg);
            pts[i0] = s0; pts[i1] = x;
        }
        __syncthreads();
    }
}

// init: FK20 table rows.  grid = 64 (vector index i), block = 64.
// S_i[m] = monomial[4031 - i - 64 m], m = 0..62, padded to 128 with the identity (fk20.go:29-33,
// 141-167; toeplitz.go:76-87).  Output: fk_pts[q*64 + i] = FFT128(S_i)[brp q], affine.
static __global__ void __launch_bounds__(64) k_fk20_table_fft(const G1Aff *__restrict__ monomial, G1 *__restrict__ fk_pts, const int8_t *__restrict__ digits) {
    __shared__ G1 pts[128];
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int m = tid; m < 128; m += 64) pts[m] = m < 63 ? G1::from_affine(monomial[4031 - i - 64 * m]) : G1::infinity();
    __syncthreads();
    g1_fft128_smem<false, false>(pts, digits, tid);
    for (int q = tid; q < 128; q += 64) fk_pts[(size_t)q * 64 + i] = pts[q];
}

}  // namespace kzg
