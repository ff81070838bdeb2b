// libkzgb200.so -- throughput kernels of the verification paths, compiled at full optimisation
// (kzgb200_verify.cu, the host side of the verifiers, is built with -Xptxas -O1; see build.py).
// Only launch wrappers are exported to the other translation units.
#include "ctx.cuh"
#include "vmsm.cuh"
#include "pairing_lanes.cuh"

#define CUL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(KZGB200_ERR_CUDA, #call, e_); } while (0)

int vm_g1_check(cudaStream_t st, const uint8_t *in48, G1Aff *out, int32_t *status, size_t n, int per_status, int subgroup) {
    if (!n) return 0;
    k_g1_check<<<(unsigned)((n + 63) / 64), 64, 0, st>>>(in48, out, status, n, per_status, subgroup);
    CUL(cudaGetLastError());
    return 0;
}

int vm_cell_coeff_digits(cudaStream_t st, const Fr &seed, const uint32_t *batch_of, const uint64_t *batch_start, const uint64_t *cell_idx,
                         const Fr *roots, Fr *rpow, int8_t *digits, size_t n) {
    if (!n) return 0;
    k_cell_coeff_digits<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, batch_of, batch_start, cell_idx, roots, rpow, digits, n);
    CUL(cudaGetLastError());
    return 0;
}

size_t vm_scratch_bytes(size_t n_items, int nw) { return n_items * (size_t)nw * KZG_VM_BUCKETS * sizeof(G1); }

// per verdict b and window w (nw windows starting at digit column w_lo): WSb[b][w] = window sum
int vm_msm_windows(cudaStream_t st, const G1Aff *points, const int8_t *digits, int TW, int w_lo, int nw,
                   const uint64_t *item_start, const uint64_t *item_end, size_t n_items, const uint64_t *batch_item_off, size_t nb,
                   G1 *scratch, G1 *WS, G1 *WSb) {
    if (n_items) {
        k_vmsm_buckets<<<(unsigned)n_items, nw, 0, st>>>(points, digits, TW, w_lo, item_start, item_end, scratch);
        const size_t n_tasks = n_items * (size_t)nw;
        k_vmsm_bucket_reduce<<<(unsigned)((n_tasks + 127) / 128), 128, 0, st>>>(scratch, WS, n_tasks);
    }
    if (nb) k_vmsm_item_reduce<<<(unsigned)nb, nw, 0, st>>>(WS, batch_item_off, WSb);
    CUL(cudaGetLastError());
    return 0;
}

int vm_combine(cudaStream_t st, const G1 *WSb, int TW, int n_segs, G1 *out, size_t nb) {
    if (!nb) return 0;
    k_vmsm_combine<<<dim3((unsigned)((nb + 31) / 32), n_segs), 32, 0, st>>>(WSb, TW, out, nb);
    CUL(cudaGetLastError());
    return 0;
}

int vm_rlc_coeff_digits(cudaStream_t st, const Fr &seed, int unit_coeff, const uint32_t *z, const uint32_t *y, const int32_t *status,
                        Fr *fy, int8_t *digits, size_t n) {
    if (!n) return 0;
    k_rlc_coeff_digits<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, unit_coeff, z, y, status, fy, digits, n);
    CUL(cudaGetLastError());
    return 0;
}

int vm_pairing_check(cudaStream_t st, const PairingConsts *pc, const G1 *A, int qa, const G1 *B, int qb, const int32_t *pre_status, int32_t *result, size_t n) {
    if (!n) return 0;
    const unsigned per_block = 128 / KZG_PL_GROUP;
    k_pairing_lanes<<<(unsigned)((n + per_block - 1) / per_block), 128, 0, st>>>(pc, A, qa, B, qb, pre_status, result, n);
    CUL(cudaGetLastError());
    return 0;
}

// Open / EvaluateLagrangePolynomial for m blobs on stream st (three launches, see kzg4844.cuh).
// vm_eval_scratch sizes the scratch for `total` blobs; a call works on scratch slots [slot, slot + m),
// so calls on different streams must use disjoint slot ranges.
int vm_eval_scratch(kzgb200_ctx *c, size_t total) {
    int rc;
    if ((rc = c->ev_cex.ensure(total * KZG_NTT_THREADS * sizeof(Fr)))) return rc;
    if ((rc = c->ev_total.ensure(total * sizeof(Fr)))) return rc;
    return c->ev_index.ensure(total * sizeof(int32_t));
}
int vm_eval_quotient(kzgb200_ctx *c, cudaStream_t st, size_t slot, const uint8_t *d_blobs, const uint32_t *z_limbs, int32_t *d_status, uint32_t *quotient,
                     uint8_t *y_out, uint32_t *y_limbs, size_t m) {
    if (!m) return 0;
    Fr inv4096; memcpy(inv4096.v, H_FR_INV4096, sizeof inv4096.v);
    Fr *cex = (Fr *)c->ev_cex.p + slot * KZG_NTT_THREADS, *tot = (Fr *)c->ev_total.p + slot;
    int32_t *idx = (int32_t *)c->ev_index.p + slot;
    k_eval_products<<<(unsigned)m, KZG_NTT_THREADS, 0, st>>>(z_limbs, c->roots, d_status, cex, tot, idx);
    k_fr_inv_batch<<<(unsigned)((m + 31) / 32), 32, 0, st>>>(tot, d_status, m);
    k_eval_finish<<<(unsigned)m, KZG_NTT_THREADS, 0, st>>>(d_blobs, z_limbs, c->roots, d_status, cex, tot, idx, quotient, y_out, y_limbs, inv4096);
    c->launches += 3;
    CUL(cudaGetLastError());
    return 0;
}
