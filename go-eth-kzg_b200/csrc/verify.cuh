// Verification kernels: single-proof checks, the EIP-4844 random-linear-combination batch check
// and the EIP-7594 cell-proof batch check.
//
// Replaces internal/kzg/kzg_verify.go:35-231 (Verify, BatchVerifyMultiPoints, fold),
// verify.go:12-145 (the per-blob driver loops, here data-parallel) and
// internal/kzg_multi/kzg_verify.go:16-105 (VerifyMultiPointKZGProofBatch).
//
// The random challenge r is drawn on the host (CSPRNG) per verdict, like the reference's
// fr.SetRandom() (internal/kzg/kzg_verify.go:136-137) -- it is NOT a Fiat-Shamir value.
#pragma once
#include "pairing.cuh"
#include "recover.cuh"
#include "fk20.cuh"
#include "vmsm.cuh"

namespace kzg {

// [k]P for a plain little-endian scalar of 4*nwin bits (nwin = 64: full 256-bit, 32: 128-bit
// randomisers), 4-bit fixed windows, Jacobian doublings.  Arguments and result are passed BY VALUE
// on purpose (see kzgb200_verify.cu): value semantics keep the data flow explicit for the compiler.
struct Scalar256 { uint32_t v[8]; };
static __device__ __noinline__ G1 g1_mul_scalar_v(G1 P, Scalar256 ks, int nwin) {
    const uint32_t *k = ks.v;
    if (P.is_inf()) return G1::infinity();
    G1J base = jac_from_xyzz(P);
    G1JT tab[15];
    tab[0] = jac_cache(base);
#pragma unroll 1
    for (int i = 1; i < 15; ++i) {
        G1J t;
        if (i & 1) { t.X = tab[i >> 1].X; t.Y = tab[i >> 1].Y; t.Z = tab[i >> 1].Z; jac_dbl(t); }     // (i+1) even: 2 * ((i+1)/2)
        else { t.X = tab[i - 1].X; t.Y = tab[i - 1].Y; t.Z = tab[i - 1].Z; jac_add(t, tab[0]); }
        tab[i] = jac_cache(t);
    }
    G1J acc; acc.X = Fp::zero(); acc.Y = Fp::zero(); acc.Z = Fp::zero();
#pragma unroll 1
    for (int w = nwin - 1; w >= 0; --w) {
        if (w != nwin - 1) {
#pragma unroll 1
            for (int i = 0; i < 4; ++i) jac_dbl(acc);
        }
        unsigned d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) jac_add(acc, tab[d - 1]);
    }
    return jac_to_xyzz(acc);
}
__device__ __forceinline__ void g1_mul_scalar(G1 *out, const G1 *pp, const uint32_t *k, int nwin = 64) {
    Scalar256 ks;
#pragma unroll
    for (int i = 0; i < 8; ++i) ks.v[i] = k[i];
    G1 r = g1_mul_scalar_v(*pp, ks, nwin);
    *out = r;
}
// base^e for a small exponent e
static __device__ __noinline__ Fr fr_pow_u64(Fr base, unsigned long long e) {
    Fr r = Fr::one();
#pragma unroll 1
    for (int bit = 63 - __clzll(e | 1ULL); bit >= 0; --bit) {
        r = fr_sqr_ni(r);
        if ((e >> bit) & 1) r = fr_mul_ni(r, base);
    }
    return r;
}

// ---- n independent single-proof checks (kzg_verify.go:35-100 rewritten to fixed G2) ------------
// status[i] must hold OK or an earlier decode error; z/y are plain limbs.
// PA[i] = -(C - [y]G + [z]pi), PB[i] = pi; the check e(PA, G2) e(PB, [s]G2) == 1 runs in k_pairing_lanes.
// Four threads per item: [y]G and [z]pi are independent, and each is GLV-split (vmsm.cuh: s = k1 + k2 * (-x^2), halves below
// 2^127) into two 128-bit scalar multiplications, so a one-item call (the reference's calling pattern: pure latency) is ~124
// dependent doublings deep instead of ~252.  Role 0: [k1(y)]G, 1: phi2([k2(y)]G), 2: [k1(z)]pi, 3: phi2([k2(z)]pi); the four
// parts meet in shared memory and role 0 forms C - [y]G + [z]pi.
static __global__ void __launch_bounds__(64) k_verify_single_prep(const G1Aff *__restrict__ commitments, const G1Aff *__restrict__ proofs,
                                                           const uint32_t *__restrict__ z, const uint32_t *__restrict__ y,
                                                           const G1Aff *__restrict__ g1_gen, const int32_t *__restrict__ status,
                                                           G1 *__restrict__ PA, G1 *__restrict__ PB, size_t n) {
    __shared__ G1 hand[16][3];
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int role = threadIdx.x & 3, slot = threadIdx.x >> 2;
    const bool live = i < n && status[i] == ST_OK;
    G1 part = G1::infinity();
    if (live) {
        uint32_t k1[4], k2[4];
        bool n1, n2;
        glv_split(((role & 2) ? z : y) + i * 8, k1, n1, k2, n2);
        Scalar256 ks;
#pragma unroll
        for (int q = 0; q < 4; ++q) { ks.v[q] = (role & 1) ? k2[q] : k1[q]; ks.v[4 + q] = 0; }
        G1 base = (role & 2) ? G1::from_affine(proofs[i]) : G1::from_affine(*g1_gen);
        part = g1_mul_scalar_v(base, ks, 32);
        if ((role & 1) ? n2 : n1) part.neg_inplace();
        if (role & 1) part = g1_phi2(part);
    }
    if (role) hand[slot][role - 1] = part;
    __syncthreads();
    if (role || i >= n) return;
    if (!live) { PA[i] = G1::infinity(); PB[i] = G1::infinity(); return; }
    G1 yG = part;
    g1_add(yG, hand[slot][0]);
    yG.neg_inplace();
    G1 A = G1::from_affine(commitments[i]);
    g1_add(A, yG);
    g1_add(A, hand[slot][1]);
    g1_add(A, hand[slot][2]);              // C - [y]G + [z]pi
    A.neg_inplace();
    PA[i] = A; PB[i] = G1::from_affine(proofs[i]);
}

// ---- EIP-4844 RLC batch (kzg_verify.go:111-231) -------------------------------------------------
// The three sums  sum r_i pi_i,  sum r_i C_i,  sum r_i z_i pi_i  are bucket MSMs (vmsm.cuh) over the
// point array [proofs | commitments] with two verdict slots; comb[seg * 2 + slot] are their 32-window
// segments.  fsum = sum r_i y_i goes through the fixed-base table of the monomial SRS (point 0 = G).
// one block: scal[0..7] = plain limbs of sum fy[i], scal[8..511] = 0  (a 64-scalar row for k_msm_fixed)
static __global__ void __launch_bounds__(128) k_rlc_fsum(const Fr *__restrict__ fy, size_t n, uint32_t *__restrict__ scal) {
    __shared__ uint32_t sf[8 * 128];
    const int tid = threadIdx.x;
    Fr f = Fr::zero();
    for (size_t i = tid; i < n; i += 128) f = Fr::add(f, ld_fr(fy + i));
    sm_store<128>(sf, tid, f);
    __syncthreads();
    for (int st = 64; st > 0; st >>= 1) {
        if (tid < st) sm_store<128>(sf, tid, Fr::add(sm_load<128>(sf, tid), sm_load<128>(sf, tid + st)));
        __syncthreads();
    }
    for (int i = tid; i < 512; i += 128) scal[i] = 0;
    __syncthreads();
    if (tid == 0) {
        Fr fsum = fr_from_mont(sm_load<128>(sf, 0));
#pragma unroll
        for (int q = 0; q < 8; ++q) scal[q] = fsum.v[q];
    }
}
// PA[0] = sum r_i C_i - [sum r_i y_i]G + sum r_i z_i pi_i,  PB[0] = -sum r_i pi_i;  the check
// e(PA, G2) e(PB, [s]G2) == 1 runs in k_pairing_lanes (qa = 0, qb = 1)   (kzg_verify.go:160-193)
static __global__ void k_rlc_prep(const G1 *__restrict__ comb, const G1 *__restrict__ yG, G1 *__restrict__ PA, G1 *__restrict__ PB) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    G1 A = comb[1];                         // seg 0, slot 1: sum r_i C_i
    G1 t = *yG; t.neg_inplace();
    g1_add(A, t);
    g1_add(A, comb[2]);                     // seg 1, slot 0: k1 half of sum r_i z_i pi_i
    g1_add(A, g1_phi2(comb[4]));            // seg 2, slot 0: k2 half
    G1 B = comb[0];                         // seg 0, slot 0: sum r_i pi_i
    B.neg_inplace();
    PA[0] = A; PB[0] = B;
}

// ---- EIP-7594 cell batch (kzg_multi/kzg_verify.go:16-105) ---------------------------------------
// status[i] = status[i] if that is an error, else later[i]
static __global__ void k_status_merge(int32_t *__restrict__ status, const int32_t *__restrict__ later, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == ST_OK) status[i] = later[i];
}
// First error of a verdict in the REFERENCE's order (api_eip7594.go:190-213: every unique commitment in row order, then every
// proof in index order, then every cell in index order): key = stage << 56 | index << 8 | status, minimum per verdict.
// stage_arg 0: status[] are the unique commitments' decode results (index = row); 1: the per-cell slots, which hold a proof's
// decode error (stage 1) or, if the proof was fine, the cell's NON_CANONICAL_SCALAR (stage 2).
static __global__ void k_merge_status(const int32_t *__restrict__ status, const uint32_t *__restrict__ group_of, unsigned long long *__restrict__ batch_key, size_t n, int stage_arg) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t st = status[i];
    if (st == ST_OK) return;
    const unsigned long long stage = stage_arg == 0 ? 0ull : (st == ST_NON_CANONICAL_SCALAR ? 2ull : 1ull);
    atomicMin(&batch_key[group_of[i]], (stage << 56) | ((unsigned long long)i << 8) | (unsigned long long)(uint32_t)st);
}
// batch_status[b] keeps a host-side error (cell index out of range: checked before any decoding, api_eip7594.go:184-188), else takes the keyed first error
static __global__ void k_status_finish(const unsigned long long *__restrict__ batch_key, int32_t *__restrict__ batch_status, size_t nb) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (batch_status[b] == ST_OK && batch_key[b] != ~0ull) batch_status[b] = (int32_t)(batch_key[b] & 0xffull);
}
// work item w = (cells [start, end) of ONE batch): partial[w][64] = sum_k r^k * interpolation poly of cell k.
// One warp per cell, 8 warps per block.  interpolation = CosetIFFT_64(brp(evals)) on the coset
// h_c <w_64>, h_c = w_8192^brp7(c)  (kzg_verify.go:51-66, srs.go:60-103, coset_fft.go:59-70)
static __global__ void __launch_bounds__(256) k_cell_interp(const uint8_t *__restrict__ cells, const uint64_t *__restrict__ cell_idx, const Fr *__restrict__ rpow,
                                                     const uint64_t *__restrict__ item_start, const uint64_t *__restrict__ item_end,
                                                     const Fr *__restrict__ roots, Fr inv64, int32_t *__restrict__ status, Fr *__restrict__ partial) {
    __shared__ uint32_t sm[8][64 * 8];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t s = item_start[blockIdx.x], e = item_end[blockIdx.x];
    Fr acc0 = Fr::zero(), acc1 = Fr::zero();
    Fr r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) r2.v[i] = FR_R2[i];
    uint32_t *buf = sm[w];
    for (uint64_t k = s + w; k < e; k += 8) {
        // load the 64 evaluations (already in bit-reversed order == DIT input order)
        int bad = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int j = lane + 32 * h;
            Fr x;
            load_be32(x.v, cells + k * 2048 + (size_t)j * 32);
            if (!fr_is_canonical(x.v)) bad = 1;
            sm_store<64>(buf, j, fr_mul_ni(x, r2));
        }
        bad = __any_sync(0xffffffffu, bad);
        __syncwarp();
        if (bad) { if (lane == 0) atomicCAS(&status[k], (int32_t)ST_OK, (int32_t)ST_NON_CANONICAL_SCALAR); continue; }
        if (status[k] != ST_OK) continue;      // proof failed to decode: the batch is an error anyway
        // DIT inverse size-64 (twiddles w_64^-t = w_8192^-(128 t))
#pragma unroll 1
        for (int st = 0; st < 6; ++st) {
            int half = 1 << st;
            int j = lane & (half - 1);
            int i0 = ((lane >> st) << (st + 1)) + j, i1 = i0 + half;
            int t = j * (32 >> st) * 128;
            Fr x = sm_load<64>(buf, i0), y = sm_load<64>(buf, i1);
            if (t) y = fr_mul_ni(y, ld_fr(roots + (ROOTS_N - t)));
            sm_store<64>(buf, i0, Fr::add(x, y));
            sm_store<64>(buf, i1, Fr::sub(x, y));
            __syncwarp();
        }
        Fr scale = fr_mul_ni(rpow[k], inv64);
        int hexp = (int)(__brev((unsigned)cell_idx[k]) >> 25);          // brp7(c): h_c = w_8192^hexp
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int j = lane + 32 * h;
            int t = (hexp * j) & (ROOTS_N - 1);
            Fr v = fr_mul_ni(sm_load<64>(buf, j), scale);
            if (t) v = fr_mul_ni(v, ld_fr(roots + (ROOTS_N - t)));       // h_c^-j
            if (h == 0) acc0 = Fr::add(acc0, v); else acc1 = Fr::add(acc1, v);
        }
        __syncwarp();
    }
    // cross-warp sum
    __syncthreads();
    sm_store<64>(sm[w], lane, acc0);
    sm_store<64>(sm[w], lane + 32, acc1);
    __syncthreads();
    if (threadIdx.x < 64) {
        Fr tot = Fr::zero();
        for (int q = 0; q < 8; ++q) tot = Fr::add(tot, sm_load<64>(sm[q], threadIdx.x));
        partial[(size_t)blockIdx.x * 64 + threadIdx.x] = tot;
    }
}
// interp[batch][j] (plain limbs, for k_msm_fixed) = sum of the batch's partials
static __global__ void k_cell_interp_reduce(const Fr *__restrict__ partial, const uint64_t *__restrict__ batch_item_off, uint32_t *__restrict__ interp) {
    const int b = blockIdx.x, j = threadIdx.x;   // 64 threads
    Fr tot = Fr::zero();
    for (uint64_t w = batch_item_off[b]; w < batch_item_off[b + 1]; ++w) tot = Fr::add(tot, partial[w * 64 + j]);
    Fr p = fr_from_mont(tot);
    for (int q = 0; q < 8; ++q) interp[((size_t)b * 64 + j) * 8 + q] = p.v[q];
}
// per (small) verdict: comm[b] = sum over its unique commitments of [w_row] C_row, w_row = sum of r_k over the row's cells (kzg_verify.go:37-45).
// One thread per verdict, a ~134-bit scalar multiplication per row: pure latency (1.9 ms), which is why it is a kernel of its own on the
// verifier's side chain instead of a step of k_cell_prep after the bucket MSM.  host_status: the statuses known before any decoding (cell
// index range); rows whose commitment failed to decode are the point at infinity and contribute nothing (the verdict is an error anyway).
static __global__ void __launch_bounds__(32) k_cell_row_weights(const G1Aff *__restrict__ uniq_commit, const uint64_t *__restrict__ row_off,
                                                           const uint64_t *__restrict__ batch_row_off, const uint32_t *__restrict__ row_cells,
                                                           const Fr *__restrict__ rpow, const int32_t *__restrict__ host_status,
                                                           const int32_t *__restrict__ large_of /*may be null*/, G1 *__restrict__ comm, size_t n_batches) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batches) return;
    G1 comms = G1::infinity();
    if (host_status[b] == ST_OK && !(large_of && large_of[b] >= 0)) {
        for (uint64_t row = batch_row_off[b]; row < batch_row_off[b + 1]; ++row) {
            Fr wsum = Fr::zero();
            for (uint64_t q = row_off[row]; q < row_off[row + 1]; ++q) wsum = Fr::add(wsum, rpow[row_cells[q]]);
            Fr wp = fr_from_mont(wsum);
            G1Aff ca = uniq_commit[row];
            if (ca.is_inf()) continue;
            G1 C = G1::from_affine(ca), t;
            int top = 0;                              // the weight is a sum of 126-bit coefficients: ~134 bits for 128 cells
#pragma unroll
            for (int q = 0; q < 8; ++q) if (wp.v[q]) top = q;
            g1_mul_scalar(&t, &C, wp.v, 8 * (top + 1));
            g1_add(comms, t);
        }
    }
    comm[b] = comms;
}
// per batch: final combination; PA[b], PB[b] feed k_pairing_lanes (qa = 2, qb = 0).  comm_small[b]: k_cell_row_weights; comm_large[lb]: the bucket MSM
// over a large verdict's commitments.
static __global__ void __launch_bounds__(32) k_cell_prep(const G1 *__restrict__ comb /*[seg][batch]: seg 0 = sum r pi, 1 and 2 = GLV halves of sum r h^64 pi*/, const G1 *__restrict__ interp_commit,
                                                    const G1 *__restrict__ comm_small, const int32_t *__restrict__ batch_status,
                                                    const int32_t *__restrict__ large_of /*may be null*/, const G1 *__restrict__ comm_large,
                                                    G1 *__restrict__ PA, G1 *__restrict__ PB, size_t n_batches) {
    size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_batches) return;
    if (batch_status[b] != ST_OK) { PA[b] = G1::infinity(); PB[b] = G1::infinity(); return; }
    G1 sumS = comb[b], sumW = comb[n_batches + b];
    g1_add(sumW, g1_phi2(comb[2 * n_batches + b]));
    const int lb = large_of ? large_of[b] : -1;
    G1 comms = lb >= 0 ? comm_large[lb] : comm_small[b];
    G1 I = interp_commit[b];
    I.neg_inplace();
    g1_add(comms, I);
    g1_add(comms, sumW);                  // rl = sum w C - [interp] + sum r^k h^64 pi   (kzg_verify.go:85-87)
    comms.neg_inplace();
    PA[b] = sumS; PB[b] = comms;          // e(sum r^k pi, [s^64]G2) e(-rl, G2) == 1
}

}  // namespace kzg
