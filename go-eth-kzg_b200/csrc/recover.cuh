// EIP-7594 erasure recovery on device: (E*Z)/Z on the coset 7*<w_8192>.
//
// Replaces api_eip7594.go:93-142 (recoverPolynomialCoeffs) and
// internal/erasure_code/erasure_code.go:75-164 (constructVanishingPolyOnIndices,
// RecoverPolynomialCoefficients, vanishingPolyCoeff).
//
// Structure exploited (the reference runs five generic size-8192 transforms):
//  * Z(x) = Zs(x^64) with deg Zs <= 64, so FFT_8192(Z) and cosetFFT_8192(Z) are 128-periodic: in
//    the bit-reversed (cell) order every scalar of cell c sees the same factor
//        zev[c]  = Zs(w_128^brp7(c))          and      czv[c] = Zs(7^64 * w_128^brp7(c)),
//    both produced by one size-128 DIF transform (output position c == cell index).
//  * size-8192 transforms are done as two shared-memory size-4096 transforms (one CTA each) plus
//    one radix-2 stage fused into the neighbouring element-wise pass; orders are chosen so that no
//    bit-reversal pass exists: cells arrive in brp order -> DIT inverse -> natural -> (coset scale)
//    -> DIF forward -> brp -> divide -> DIT inverse -> natural coefficients.
#pragma once
#include "kzg4844.cuh"

namespace kzg {

// pow7[k] = 7^k (k < 8192) and ipow7[k] = 7^-k, Montgomery; one thread per k
static __global__ void k_init_pow7(Fr *pow7, Fr *ipow7, Fr seven, Fr inv7) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 8192) return;
    Fr a = Fr::one(), b = Fr::one();
    for (int bit = 12; bit >= 0; --bit) {
        a = fr_sqr_ni(a); b = fr_sqr_ni(b);
        if ((t >> bit) & 1) { a = fr_mul_ni(a, seven); b = fr_mul_ni(b, inv7); }
    }
    pow7[t] = a; ipow7[t] = b;
}

// K0: per blob (128 threads): vanishing polynomial of the missing cells and its two evaluation
// tables.  present[blob][c] != 0 marks provided cells.  Outputs zev[blob][128], czinv[blob][128].
static __global__ void __launch_bounds__(128) k_rec_vanishing(const uint8_t *__restrict__ present, const int32_t *__restrict__ status,
                                                       const Fr *__restrict__ roots, const Fr *__restrict__ pow7,
                                                       Fr *__restrict__ zev, Fr *__restrict__ czinv) {
    __shared__ uint32_t za[128 * 8];   // Zs coefficients (limb planes), later FFT workspace
    __shared__ uint32_t zb[128 * 8];
    __shared__ int miss[128];
    __shared__ int n_miss;
    const int blob = blockIdx.x, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    if (tid == 0) {
        int n = 0;
        for (int c = 0; c < 128; ++c) if (!present[(size_t)blob * 128 + c]) miss[n++] = (int)(__brev((unsigned)c) >> 25);   // brp_128(cellID), api_eip7594.go:117-122
        n_miss = n;
    }
    // Zs = 1
    sm_store<128>(za, tid, tid == 0 ? Fr::one() : Fr::zero());
    __syncthreads();
    const int nm = n_miss;
    // multiply by (x - w_128^m) one root at a time: c_i <- c_{i-1} - r*c_i   (erasure_code.go:151-164)
    for (int k = 0; k < nm; ++k) {
        Fr r = ld_fr(roots + miss[k] * 64);
        Fr ci = sm_load<128>(za, tid);
        Fr cim1 = tid ? sm_load<128>(za, tid - 1) : Fr::zero();
        __syncthreads();
        sm_store<128>(za, tid, Fr::sub(cim1, fr_mul_ni(r, ci)));
        __syncthreads();
    }
    // coset copy: coefficient i scaled by (7^64)^i = pow7[64 i]
    Fr ci = sm_load<128>(za, tid);
    sm_store<128>(zb, tid, tid ? fr_mul_ni(ci, ld_fr(pow7 + 64 * tid)) : ci);
    __syncthreads();
    // size-128 DIF transforms need 64 butterflies per stage: run both with all 128 threads, 64 each
    uint32_t *mine = tid < 64 ? za : zb;
    {
        const int t64 = tid & 63;
#pragma unroll 1
        for (int s = 0; s < 7; ++s) {
            const int log_half = 6 - s, half = 1 << log_half;
            int j = t64 & (half - 1);
            int i0 = ((t64 >> log_half) << (log_half + 1)) + j, i1 = i0 + half;
            int t = j * (64 >> log_half) * 64;
            Fr x = sm_load<128>(mine, i0), y = sm_load<128>(mine, i1);
            Fr d = Fr::sub(x, y);
            if (t) d = fr_mul_ni(d, ld_fr(roots + t));
            sm_store<128>(mine, i0, Fr::add(x, y));
            sm_store<128>(mine, i1, d);
            __syncthreads();
        }
    }
    zev[(size_t)blob * 128 + tid] = sm_load<128>(za, tid);
    czinv[(size_t)blob * 128 + tid] = fr_inv(sm_load<128>(zb, tid));   // never zero: the coset avoids the domain
}

// K1: ez[blob][i] = E_brp[i] * zev[cell(i)]  (zero for missing cells); canonical check of the cells.
// cell_slot[blob][c] = index of cell c inside this blob's provided cells (-1 if missing); cells_base[blob] = first cell.
static __global__ void k_rec_scale(const uint8_t *__restrict__ cells, const int32_t *__restrict__ cell_slot, const uint64_t *__restrict__ cells_base,
                            const Fr *__restrict__ zev, int32_t *__restrict__ status, Fr *__restrict__ ez) {
    const int blob = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0..8191
    if (status[blob] != ST_OK && status[blob] != ST_NON_CANONICAL_SCALAR) return;
    const int c = i >> 6;
    int slot = cell_slot[(size_t)blob * 128 + c];
    Fr v = Fr::zero();
    if (slot >= 0) {
        Fr x;
        load_be32(x.v, cells + (cells_base[blob] + (uint64_t)slot) * 2048 + (size_t)(i & 63) * 32);
        if (!fr_is_canonical(x.v)) atomicMax(&status[blob], (int32_t)ST_NON_CANONICAL_SCALAR);
        Fr r2;
#pragma unroll
        for (int k = 0; k < 8; ++k) r2.v[k] = FR_R2[k];
        v = fr_mul_ni(fr_mul_ni(x, r2), ld_fr(zev + (size_t)blob * 128 + c));
    }
    st_fr(ez + (size_t)blob * 8192 + i, v);
}

// size-4096 inverse DIT on one half of a brp-ordered size-8192 input.  grid = (2, blobs).
// out half h, natural order, unscaled:  h=0 -> A[k] (even samples), h=1 -> B[k] (odd samples)
static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_half_dit_inv(const Fr *__restrict__ in, Fr *__restrict__ out, const int32_t *__restrict__ status,
                                                                  const Fr *__restrict__ roots) {
    extern __shared__ uint32_t sm[];
    const int h = blockIdx.x, blob = blockIdx.y, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    const Fr *src = in + (size_t)blob * 8192 + h * 4096;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) sm_store<4096>(sm, i, ld_fr(src + i));
    __syncthreads();
    ntt_smem<12, true, true>(sm, roots, tid, KZG_NTT_THREADS);
    Fr *dst = out + (size_t)blob * 8192 + h * 4096;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) st_fr(dst + i, sm_load<4096>(sm, i));
}

// final radix-2 stage of the inverse size-8192 DIT fused with 1/8192 and a per-index scale table:
//   X[k] = (A[k] + w^-k B[k]) / 8192 * scale[k],  X[k+4096] = (A[k] - w^-k B[k]) / 8192 * scale[k+4096]
// upper == 0: only k < 4096 is produced and written densely to out[blob][4096] (the coefficients).
static __global__ void k_inv_combine(const Fr *__restrict__ ab, Fr *__restrict__ out, const int32_t *__restrict__ status, const Fr *__restrict__ roots,
                              const Fr *__restrict__ scale, Fr inv_n, int upper) {
    const int blob = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0..4095
    if (status[blob] != ST_OK) return;
    const Fr *A = ab + (size_t)blob * 8192, *B = A + 4096;
    Fr a = ld_fr(A + k), b = ld_fr(B + k);
    if (k) b = fr_mul_ni(b, ld_fr(roots + (ROOTS_N - k)));
    Fr lo = fr_mul_ni(fr_mul_ni(Fr::add(a, b), inv_n), ld_fr(scale + k));
    if (upper) {
        Fr hi = fr_mul_ni(fr_mul_ni(Fr::sub(a, b), inv_n), ld_fr(scale + k + 4096));
        st_fr(out + (size_t)blob * 8192 + k, lo);
        st_fr(out + (size_t)blob * 8192 + k + 4096, hi);
    } else {
        st_fr(out + (size_t)blob * 4096 + k, lo);
    }
}

// forward size-8192 DIF: first radix-2 stage on load, then a size-4096 DIF in shared memory;
// output (brp order) multiplied by percell[blob][pos >> 6].  grid = (2, blobs).
static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_half_dif_fwd(const Fr *__restrict__ in, Fr *__restrict__ out, const int32_t *__restrict__ status,
                                                                  const Fr *__restrict__ roots, const Fr *__restrict__ percell) {
    extern __shared__ uint32_t sm[];
    const int h = blockIdx.x, blob = blockIdx.y, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    const Fr *src = in + (size_t)blob * 8192;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x = ld_fr(src + i), y = ld_fr(src + i + 4096);
        Fr v;
        if (h == 0) v = Fr::add(x, y);
        else { v = Fr::sub(x, y); if (i) v = fr_mul_ni(v, ld_fr(roots + i)); }
        sm_store<4096>(sm, i, v);
    }
    __syncthreads();
    ntt_smem<12, false, false>(sm, roots, tid, KZG_NTT_THREADS);
    Fr *dst = out + (size_t)blob * 8192 + h * 4096;
    const Fr *pc = percell + (size_t)blob * 128 + h * 64;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) st_fr(dst + i, fr_mul_ni(sm_load<4096>(sm, i), ld_fr(pc + (i >> 6))));
}

// all 128 cells from monomial coefficients (fk20.go:58-74): cells 0..63 = brp(FFT_4096(c)),
// cells 64..127 = brp(FFT_4096(c_j w_8192^j)).  grid = (2, blobs).
static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_cells_from_coeffs(const Fr *__restrict__ coeffs, uint8_t *__restrict__ cells,
                                                                       const int32_t *__restrict__ status, const Fr *__restrict__ roots) {
    extern __shared__ uint32_t sm[];
    const int h = blockIdx.x, blob = blockIdx.y, tid = threadIdx.x;
    uint8_t *dst = cells + (size_t)blob * 262144 + (size_t)h * 131072;
    if (status[blob] != ST_OK) {
        uint4 *o = reinterpret_cast<uint4 *>(dst);
        for (int i = tid; i < 131072 / 16; i += KZG_NTT_THREADS) o[i] = make_uint4(0, 0, 0, 0);
        return;
    }
    const Fr *src = coeffs + (size_t)blob * 4096;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x = ld_fr(src + i);
        if (h && i) x = fr_mul_ni(x, ld_fr(roots + i));
        sm_store<4096>(sm, i, x);
    }
    __syncthreads();
    ntt_smem<12, false, false>(sm, roots, tid, KZG_NTT_THREADS);
    Fr one_plain = Fr::zero(); one_plain.v[0] = 1;
    for (int i = tid; i < 4096; i += KZG_NTT_THREADS) {
        Fr x = fr_mul_ni(sm_load<4096>(sm, i), one_plain);
        store_be32(dst + i * 32, x.v);
    }
}

}  // namespace kzg
