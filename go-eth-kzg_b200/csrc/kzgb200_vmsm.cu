// libkzgb200.so -- throughput kernels of the verification paths, compiled at full optimisation
// (kzgb200_verify.cu, the host side of the verifiers, is built with -Xptxas -O1; see build.py).
// Only launch wrappers are exported to the other translation units.
#include "ctx.cuh"
#include "vmsm.cuh"
#include "pairing_lanes.cuh"
#include "pairing_lanes8.cuh"

#define CUL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(KZGB200_ERR_CUDA, #call, e_); } while (0)

int vm_g1_check(cudaStream_t st, const uint8_t *in48, G1Aff *out, int32_t *status, size_t n, int per_status, int subgroup) {
    if (!n) return 0;
    // tunables "decode_dual" (paired squarings in the subgroup test) and "decode_minb" (resident CTAs of 64 threads per SM the kernel is compiled
    // for: 4 -> 255 registers, 6 -> 168, 8 -> 128)
    const unsigned grid = (unsigned)((n + 63) / 64);
#define KZG_DECODE(M, B) k_g1_check<M, B><<<grid, 64, 0, st>>>(in48, out, status, n, per_status, subgroup)
    if (g_decode_dual) { if (g_decode_minb == 8) KZG_DECODE(MulCall2, 8); else if (g_decode_minb == 6) KZG_DECODE(MulCall2, 6); else KZG_DECODE(MulCall2, 4); }
    else { if (g_decode_minb == 8) KZG_DECODE(MulCall, 8); else if (g_decode_minb == 6) KZG_DECODE(MulCall, 6); else KZG_DECODE(MulCall, 4); }
#undef KZG_DECODE
    CUL(cudaGetLastError());
    return 0;
}

int vm_cell_coeff_digits(cudaStream_t st, const Fr &seed, const uint32_t *batch_of, const uint64_t *batch_start, const uint64_t *cell_idx,
                         const Fr *roots, Fr *rpow, int8_t *digits, int8_t *digits256, size_t n) {
    if (!n) return 0;
    k_cell_coeff_digits<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, batch_of, batch_start, cell_idx, roots, rpow, digits, digits256, n);
    CUL(cudaGetLastError());
    return 0;
}

size_t vm_scratch_bytes(size_t n_items, int nw, int nbuckets) { return n_items * (size_t)nw * nbuckets * sizeof(G1); }

// per verdict b and window w (nw windows starting at digit column w_lo): WSb[b][w] = window sum.  nbuckets = 8 (signed 4-bit
// digits) or 128 (signed 8-bit digits); order (optional): item positions index this list of point indices.
int vm_msm_windows(cudaStream_t st, const G1Aff *points, const int8_t *digits, int TW, int w_lo, int nw, const uint32_t *order, int nbuckets,
                   const uint64_t *item_start, const uint64_t *item_end, size_t n_items, const uint64_t *batch_item_off, size_t nb,
                   G1 *scratch, G1 *WS, G1 *WSb) {
    if (n_items) {
        const size_t n_tasks = n_items * (size_t)nw;
#define KZG_VMSM_LAUNCH(NB, POL) k_vmsm_buckets<NB, POL><<<(unsigned)n_items, nw, 0, st>>>(points, digits, TW, w_lo, order, item_start, item_end, scratch)
#define KZG_VMSM_PICK(NB) do { switch (g_vmsm_policy) { case 0: KZG_VMSM_LAUNCH(NB, MulInline); break; case 2: KZG_VMSM_LAUNCH(NB, MulInlineLazy); break; \
                                                        case 3: KZG_VMSM_LAUNCH(NB, MulCall); break; default: KZG_VMSM_LAUNCH(NB, MulCallLazy); } } while (0)
        if (nbuckets == KZG_LARGE_BUCKETS) {
            KZG_VMSM_PICK(KZG_LARGE_BUCKETS);
            k_vmsm_bucket_reduce<KZG_LARGE_BUCKETS><<<(unsigned)((n_tasks + 127) / 128), 128, 0, st>>>(scratch, WS, n_tasks);
        } else {
            KZG_VMSM_PICK(KZG_VM_BUCKETS);
            k_vmsm_bucket_reduce<KZG_VM_BUCKETS><<<(unsigned)((n_tasks + 127) / 128), 128, 0, st>>>(scratch, WS, n_tasks);
        }
#undef KZG_VMSM_PICK
#undef KZG_VMSM_LAUNCH
    }
    if (nb) {
        // few verdicts of many items each: four partial sums per window (see k_vmsm_item_reduce)
        const int par = (n_items >= 8 * nb) ? (nw <= 64 ? 4 : 2) : 1;      // <= 256 threads (the full addition needs ~200 registers)
        const size_t smem = par > 1 ? (size_t)nw * par * sizeof(G1) : 0;
        if (smem > 48 * 1024) CUL(cudaFuncSetAttribute(k_vmsm_item_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: set at every use
        k_vmsm_item_reduce<<<(unsigned)nb, dim3(nw, par), smem, st>>>(WS, batch_item_off, WSb);
    }
    CUL(cudaGetLastError());
    return 0;
}

int vm_combine(cudaStream_t st, const G1 *WSb, int TW, int nw, int dbl, int n_segs, G1 *out, size_t nb) {
    if (!nb) return 0;
    k_vmsm_combine<<<dim3((unsigned)((nb + 31) / 32), n_segs), 32, 0, st>>>(WSb, TW, nw, dbl, out, nb);
    CUL(cudaGetLastError());
    return 0;
}

int vm_row_weight_digits(cudaStream_t st, const Fr *rpow, const uint64_t *row_off, const uint32_t *row_cells, int8_t *digits, size_t n_rows) {
    if (!n_rows) return 0;
    k_row_weight_digits<<<(unsigned)((n_rows + 127) / 128), 128, 0, st>>>(rpow, row_off, row_cells, digits, n_rows);
    CUL(cudaGetLastError());
    return 0;
}

int vm_cell_columns_large(cudaStream_t st, const G1 *colsum, const uint32_t *large_ids, size_t n_large, const int8_t *glv_tw_digits, G1 *comb, size_t nb) {
    if (!n_large) return 0;
    k_cell_columns_large<<<(unsigned)n_large, 128, 0, st>>>(colsum, large_ids, glv_tw_digits, comb, nb);
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

// One pairing-product check per item.  Two forms of the same arithmetic: pl32 (one check per warp, ~1.6 k dependent Fp products:
// 2.9 ms for one check) while every check gets a resident warp, pl8 (8 lanes per check, ~4.6 k dependent products but a third of the
// issued work: 7.1 ms for 4096 checks against 15.2 ms) beyond that.  g_pairing_lanes (tunable "pairing_lanes": 0 auto, 8, 32) overrides.
int vm_pairing_check(cudaStream_t st, const PairingConsts *pc, const G1 *A, int qa, const G1 *B, int qb, const int32_t *pre_status, int32_t *result, size_t n) {
    if (!n) return 0;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    const bool wide = g_pairing_lanes == 32 || (g_pairing_lanes != 8 && n <= (size_t)sms * KZG_PL32_RESIDENT_WARPS);
    if (wide) pl32::k_pairing_lanes<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(pc, A, qa, B, qb, pre_status, result, n);
    else pl8::k_pairing_lanes<<<(unsigned)((n + 15) / 16), 128, 0, st>>>(pc, A, qa, B, qb, pre_status, result, n);
    CUL(cudaGetLastError());
    return 0;
}

// Open / EvaluateLagrangePolynomial for m blobs on stream st (three launches, see kzg4844.cuh).
// vm_eval_scratch sizes the scratch for `total` blobs; a call works on scratch slots [slot, slot + m),
// so calls on different streams must use disjoint slot ranges.
int vm_eval_scratch(kzg_lane *c, size_t total) {
    int rc;
    if ((rc = c->ev_cex.ensure(total * KZG_NTT_THREADS * sizeof(Fr)))) return rc;
    if ((rc = c->ev_total.ensure(total * sizeof(Fr)))) return rc;
    return c->ev_index.ensure(total * sizeof(int32_t));
}
int vm_eval_quotient(kzg_lane *c, cudaStream_t st, size_t slot, const uint8_t *d_blobs, const uint32_t *z_limbs, int32_t *d_status, uint32_t *quotient,
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

// ---- unit-test hooks (include/kzgb200_debug.h) -----------------------------------------------------
namespace kzg {
static __global__ void k_dbg_glv_digits(const uint32_t *__restrict__ s, int8_t *__restrict__ digits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k1[4], k2[4];
    bool n1, n2;
    glv_split(s + i * 8, k1, n1, k2, n2);
    recode16_signed(digits + i * 64, k1, n1);
    recode16_signed(digits + i * 64 + 32, k2, n2);
}
// digit rows [n][96] for a plain MSM: windows 0..31 unused (zero), 32..95 = GLV halves of s_i
static __global__ void k_dbg_vmsm_digits(const uint32_t *__restrict__ s, int8_t *__restrict__ digits, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k1[4], k2[4];
    bool n1, n2;
    glv_split(s + i * 8, k1, n1, k2, n2);
    int8_t *d = digits + (size_t)i * KZG_CELL_TW;
    for (int q = 0; q < 32; ++q) d[q] = 0;
    recode16_signed(d + 32, k1, n1);
    recode16_signed(d + 64, k2, n2);
}
static __global__ void k_dbg_vmsm_finish(const uint8_t *__restrict__ p48, G1Aff *__restrict__ pts, int n, const G1 *__restrict__ comb, uint8_t *__restrict__ out48, int phase) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 0) { if (i < n) g1_decompress(pts[i], p48 + (size_t)i * 48); return; }
    if (i) return;
    G1 a = comb[1];
    g1_add(a, g1_phi2(comb[2]));
    g1_compress(out48, g1_to_affine(a));
}
}  // namespace kzg

extern "C" int kzgb200_dbg_glv_digits(const uint32_t *s, int8_t *digits, int n) {
    uint32_t *ds; int8_t *dd;
    CUL(cudaMalloc(&ds, (size_t)n * 32)); CUL(cudaMalloc(&dd, (size_t)n * 64));
    CUL(cudaMemcpy(ds, s, (size_t)n * 32, cudaMemcpyHostToDevice));
    k_dbg_glv_digits<<<(n + 63) / 64, 64>>>(ds, dd, n);
    CUL(cudaGetLastError());
    CUL(cudaMemcpy(digits, dd, (size_t)n * 64, cudaMemcpyDeviceToHost));
    cudaFree(ds); cudaFree(dd);
    return 0;
}

extern "C" int lane_dbg_vmsm(kzg_lane *c, const uint8_t *p48, const uint32_t *s, int n, uint8_t *out48) {
    if (!c || n <= 0) return set_err(KZGB200_ERR_ARGS, "bad argument");
    CUL(cudaSetDevice(c->device));
    const size_t n_items = ((size_t)n + 127) / 128;
    std::vector<uint64_t> is(n_items), ie(n_items), off = {0, n_items};
    for (size_t k = 0; k < n_items; ++k) { is[k] = k * 128; ie[k] = std::min<uint64_t>(n, k * 128 + 128); }
    uint8_t *dp, *dout; uint32_t *ds; int8_t *dd; G1Aff *pts; G1 *scratch, *ws, *wsb, *comb; uint64_t *meta;
    CUL(cudaMalloc(&dp, (size_t)n * 48)); CUL(cudaMalloc(&ds, (size_t)n * 32)); CUL(cudaMalloc(&dd, (size_t)n * KZG_CELL_TW)); CUL(cudaMalloc(&pts, (size_t)n * sizeof(G1Aff)));
    CUL(cudaMalloc(&scratch, vm_scratch_bytes(n_items, KZG_CELL_TW, KZG_VM_BUCKETS))); CUL(cudaMalloc(&ws, n_items * KZG_CELL_TW * sizeof(G1)));
    CUL(cudaMalloc(&wsb, KZG_CELL_TW * sizeof(G1))); CUL(cudaMalloc(&comb, KZG_VM_SEGS * sizeof(G1))); CUL(cudaMalloc(&meta, (2 * n_items + 2) * 8)); CUL(cudaMalloc(&dout, 48));
    CUL(cudaMemcpy(dp, p48, (size_t)n * 48, cudaMemcpyHostToDevice)); CUL(cudaMemcpy(ds, s, (size_t)n * 32, cudaMemcpyHostToDevice));
    CUL(cudaMemcpy(meta, is.data(), n_items * 8, cudaMemcpyHostToDevice)); CUL(cudaMemcpy(meta + n_items, ie.data(), n_items * 8, cudaMemcpyHostToDevice));
    CUL(cudaMemcpy(meta + 2 * n_items, off.data(), 16, cudaMemcpyHostToDevice));
    k_dbg_vmsm_finish<<<(n + 63) / 64, 64>>>(dp, pts, n, nullptr, nullptr, 0);
    k_dbg_vmsm_digits<<<(n + 63) / 64, 64>>>(ds, dd, n);
    int rc = vm_msm_windows(nullptr, pts, dd, KZG_CELL_TW, 0, KZG_CELL_TW, nullptr, KZG_VM_BUCKETS, meta, meta + n_items, n_items, meta + 2 * n_items, 1, scratch, ws, wsb);
    if (!rc) rc = vm_combine(nullptr, wsb, KZG_CELL_TW, 32, 4, KZG_VM_SEGS, comb, 1);
    if (rc) return rc;
    k_dbg_vmsm_finish<<<1, 32>>>(nullptr, nullptr, 0, comb, dout, 1);
    CUL(cudaGetLastError());
    CUL(cudaMemcpy(out48, dout, 48, cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(ds); cudaFree(dd); cudaFree(pts); cudaFree(scratch); cudaFree(ws); cudaFree(wsb); cudaFree(comb); cudaFree(meta); cudaFree(dout);
    return 0;
}
