// libkzgb200.so -- verification entry points of the C ABI (include/kzgb200.h).
//
// Host side of the verifiers (staging, bookkeeping, launch order) plus their small preparation kernels
// (verify.cuh).  This translation unit is compiled with `-Xptxas -O1`: at ptxas' default level nvcc 12.9
// miscompiled the first version of k_verify_single; the guard is kept because
// everything launched from here directly is latency-bound one-thread-per-item code.  The throughput
// kernels of the verifiers (point decoding, bucket MSMs, the 8-lane pairing, the evaluation kernels)
// live in kzgb200_vmsm.cu at full optimisation and are reached through the vm_* launch wrappers.
#include "ctx.cuh"
#include "verify.cuh"
#include "cell_plan.hpp"
#include <chrono>
#include <cstdio>

namespace kzg {
static __global__ void k_dbg_g1_mul(const uint8_t *p48, const uint8_t *s32, uint8_t *out48, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Aff a; g1_decompress(a, p48 + i * 48);
    uint32_t k[8]; load_be32(k, s32 + i * 32);
    G1 P = G1::from_affine(a), r;
    g1_mul_scalar(&r, &P, k);
    g1_compress(out48 + i * 48, g1_to_affine(r));
}
static __global__ void k_dbg_decode_pair(const uint8_t *a48, const uint8_t *b48, G1 *A, G1 *B, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Aff a, b;
    g1_decompress(a, a48 + i * 48); g1_decompress(b, b48 + i * 48);
    A[i] = G1::from_affine(a); B[i] = G1::from_affine(b);
}
static __global__ void k_dbg_dump_pairing(const PairingConsts *pc, uint32_t *out) {
    // gamma[1] (24 limbs, plain), then q[0].A[0], q[0].B[0], q[0].A[67] (24 limbs each)
    Fp o1 = Fp::zero(); o1.v[0] = 1;
    const Fp2 *src[4] = {&pc->gamma[1], &pc->q[0].A[0], &pc->q[0].B[0], &pc->q[0].A[67]};
    for (int s = 0; s < 4; ++s) {
        Fp a = fp_mul_ni(src[s]->c0, o1), b = fp_mul_ni(src[s]->c1, o1);
        for (int k = 0; k < 12; ++k) { out[s * 24 + k] = a.v[k]; out[s * 24 + 12 + k] = b.v[k]; }
    }
}
}  // namespace kzg

// KZGB200_TRACE=1: host wall-clock (ms since the first call) at a few points of the verifiers, to stderr
static inline void trace_pt(const char *what) {
    static const bool on = [] { const char *e = getenv("KZGB200_TRACE"); return e && *e == '1'; }();
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    fprintf(stderr, "[kzgb200 trace] %9.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), what);
}

extern "C" {

// -------------------------------------------------------------------------------------------
// Verifiers (verify.go:12-169, api_eip7594.go:163-265)
// -------------------------------------------------------------------------------------------
static const size_t VERIFY_CHUNK = 4096;

// shared front end: decode commitments/proofs, obtain z and y (given, or Fiat-Shamir + evaluation)
//   blobs == nullptr: z32/y32 given (VerifyKZGProof); else z = challenge, y = p(z)
// writes: out_cm = commitments, out_pf = proofs (affine), zl = z limbs, yl = y limbs, d_status filled.
// The hashing + evaluation of the blobs runs on side streams while the main stream decodes the points
// (the SHA-256 of a blob is one thread's 2050 sequential compressions: latency, not throughput).  Host blobs
// travel in pieces on the copy stream and each piece gets a side stream of its own as soon as it has landed
// (pieces queued on ONE stream would serialise that latency); device blobs are one piece.
#define VERIFY_MAX_PIECES (KZG_G1FFT_MAX_SPLIT - 1)
static int verify_front(kzg_lane *c, const uint8_t *blobs, const uint8_t *cm48, const uint8_t *z32, const uint8_t *y32, const uint8_t *pf48,
                        size_t m, int32_t *d_status, G1Aff *out_cm, G1Aff *out_pf, uint32_t *zl, uint32_t *yl) {
    int rc;
    const void *d_cm, *d_pf, *d_z = nullptr, *d_y = nullptr;
    trace_pt("verify_front: begin");
    const bool blobs_host = blobs && !is_device_ptr(blobs);
    // With host blobs, the small inputs travel on the COPY stream ahead of the blob pieces and the main stream waits for them through an
    // event: queued on the main stream they were observed to complete only after the 0.5 GB of blob pieces queued behind them on the copy
    // stream (KZGB200_TRACE), which held back the event every side stream waits for -- no hashing under the H2D at all.
    cudaStream_t s_small = blobs_host ? c->copy_stream : c->stream;
    auto stage_small = [&](const void *user, size_t bytes, DevBuf &buf, const void **dev) -> int {
        if (bytes == 0) { *dev = buf.p; return 0; }
        if (is_device_ptr(user)) { *dev = user; return 0; }
        int r = buf.ensure(bytes);
        if (r) return r;
        CU(cudaMemcpyAsync(buf.p, user, bytes, cudaMemcpyHostToDevice, s_small));
        *dev = buf.p;
        return 0;
    };
    if ((rc = stage_small(cm48, m * 48, c->in_small, &d_cm))) return rc;
    if ((rc = stage_small(pf48, m * 48, c->in_small2, &d_pf))) return rc;
    if (blobs_host) { CU(cudaEventRecord(c->ev0, c->copy_stream)); CU(cudaStreamWaitEvent(c->stream, c->ev0, 0)); }
    const size_t piece = blobs_host ? ((m + VERIFY_MAX_PIECES - 1) / VERIFY_MAX_PIECES + 31) & ~(size_t)31 : m;
    const size_t n_pieces = blobs ? (m + piece - 1) / piece : 0;
    if (blobs_host) {
        if ((rc = c->in_bytes.ensure(m * KZGB200_BYTES_PER_BLOB))) return rc;
        for (size_t p = 0; p < n_pieces; ++p) {
            size_t po = p * piece, pm = std::min(piece, m - po);
            CU(cudaMemcpyAsync((char *)c->in_bytes.p + po * KZGB200_BYTES_PER_BLOB, blobs + po * KZGB200_BYTES_PER_BLOB, pm * KZGB200_BYTES_PER_BLOB,
                               cudaMemcpyHostToDevice, c->copy_stream));
            CU(cudaEventRecord(c->ev_piece[p], c->copy_stream));
        }
    }
    trace_pt("verify_front: blob pieces queued");
    if (!blobs) {
        if ((rc = stage_in(c, z32, m * 32, c->v_in2, &d_z))) return rc;
        if ((rc = stage_in(c, y32, m * 32, c->v_in3, &d_y))) return rc;
    }
    unsigned gb = (unsigned)((m + 63) / 64);
    c->mark(KZGB200_KC_VERIFY);
    CU(cudaMemsetAsync(d_status, 0, m * sizeof(int32_t), c->stream));
    if (!blobs) {
        k_scalars_from_be<<<gb, 64, 0, c->stream>>>((const uint8_t *)d_y, yl, d_status, m);
        k_scalars_from_be<<<gb, 64, 0, c->stream>>>((const uint8_t *)d_z, zl, d_status, m);
        c->launches += 2;
    }
    int32_t *d_blob_status = nullptr;      // the side streams report into their own array: merged below in the reference's order
    if (blobs) {
        if ((rc = vm_eval_scratch(c, m))) return rc;
        if ((rc = c->v_st3.ensure(m * sizeof(int32_t)))) return rc;
        d_blob_status = (int32_t *)c->v_st3.p;
        CU(cudaMemsetAsync(d_blob_status, 0, m * sizeof(int32_t), c->stream));
        // side streams start once the commitments are on the device and the status array is cleared
        CU(cudaEventRecord(c->ev_fork, c->stream));
        const uint8_t *d_blobs = blobs_host ? (const uint8_t *)c->in_bytes.p : blobs;
        for (size_t p = 0; p < n_pieces; ++p) {
            size_t po = p * piece, pm = std::min(piece, m - po);
            cudaStream_t sp = c->fft_streams[p];
            CU(cudaStreamWaitEvent(sp, c->ev_fork, 0));
            if (blobs_host) CU(cudaStreamWaitEvent(sp, c->ev_piece[p], 0));
            const uint8_t *pb = d_blobs + po * KZGB200_BYTES_PER_BLOB;
            launch_fiat_shamir(sp, pb, (const uint8_t *)d_cm + po * 48, zl + po * 8, pm);
            if ((rc = vm_eval_quotient(c, sp, po, pb, zl + po * 8, d_blob_status + po, nullptr, nullptr, yl + po * 8, pm))) return rc;
            CU(cudaEventRecord(c->ev_join[p], sp));
            c->launches += 1;
        }
    }
    c->mark(KZGB200_KC_DECODE);
    // commitments on the main stream, proofs on the aux stream: a decode is ~1.4 ms of ONE thread's latency per point and even 4096 + 4096
    // points under-fill the GPU, so the two launches run side by side.  The proofs report into a status array of their own, merged after
    // the commitments' (the reference decodes the commitment first: verify.go:12-41, 102-119).
    if ((rc = c->v_pst.ensure(m * sizeof(int32_t)))) return rc;
    int32_t *d_pst = (int32_t *)c->v_pst.p;
    CU(cudaEventRecord(c->ev_aux_fork, c->stream));
    CU(cudaStreamWaitEvent(c->aux_stream, c->ev_aux_fork, 0));
    CU(cudaMemsetAsync(d_pst, 0, m * sizeof(int32_t), c->aux_stream));
    if ((rc = vm_g1_check(c->aux_stream, (const uint8_t *)d_pf, out_pf, d_pst, m, 1, 1))) return rc;
    CU(cudaEventRecord(c->ev_aux_join, c->aux_stream));
    if ((rc = vm_g1_check(c->stream, (const uint8_t *)d_cm, out_cm, d_status, m, 1, 1))) return rc;
    CU(cudaStreamWaitEvent(c->stream, c->ev_aux_join, 0));
    k_status_merge<<<gb, 64, 0, c->stream>>>(d_status, d_pst, m);
    c->launches += 3;
    c->mark(KZGB200_KC_FR);          // what is left of the side streams' hashing / evaluation after the decode
    if (blobs) {
        for (size_t p = 0; p < n_pieces; ++p) CU(cudaStreamWaitEvent(c->stream, c->ev_join[p], 0));
        k_status_merge<<<gb, 64, 0, c->stream>>>(d_status, d_blob_status, m);      // commitment / proof errors come first (verify.go:102-119)
        c->launches += 1;
    }
    trace_pt("verify_front: all queued");
    return 0;
}

static int verify_independent(kzg_lane *c, const uint8_t *blobs, const uint8_t *cm48, const uint8_t *z32, const uint8_t *y32,
                              const uint8_t *pf48, size_t n, int32_t *status) {
    if (!c || (n && (!cm48 || !pf48 || !status || (!blobs && (!z32 || !y32))))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    if (n == 0) return KZGB200_OK;
    const bool st_dev = is_device_ptr(status);
    const size_t chunk = std::min(n, VERIFY_CHUNK);
    int rc;
    if ((rc = c->status.ensure(chunk * sizeof(int32_t)))) return rc;
    if ((rc = c->v_aff1.ensure(chunk * sizeof(G1Aff)))) return rc;
    if ((rc = c->v_aff2.ensure(chunk * sizeof(G1Aff)))) return rc;
    if ((rc = c->zbuf.ensure(chunk * 32))) return rc;
    if ((rc = c->ybuf.ensure(chunk * 32))) return rc;
    if ((rc = c->v_S.ensure(chunk * sizeof(G1)))) return rc;
    if ((rc = c->v_W.ensure(chunk * sizeof(G1)))) return rc;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        int32_t *d_status = st_dev ? status + off : (int32_t *)c->status.p;
        if ((rc = verify_front(c, blobs ? blobs + off * KZGB200_BYTES_PER_BLOB : nullptr, cm48 + off * 48, z32 ? z32 + off * 32 : nullptr,
                               y32 ? y32 + off * 32 : nullptr, pf48 + off * 48, m, d_status, (G1Aff *)c->v_aff1.p, (G1Aff *)c->v_aff2.p,
                               (uint32_t *)c->zbuf.p, (uint32_t *)c->ybuf.p))) return rc;
        c->mark(KZGB200_KC_VERIFY);
        k_verify_single_prep<<<(unsigned)((4 * m + 63) / 64), 64, 0, c->stream>>>((const G1Aff *)c->v_aff1.p, (const G1Aff *)c->v_aff2.p, (const uint32_t *)c->zbuf.p,
                                                                               (const uint32_t *)c->ybuf.p, c->g1_monomial, d_status, (G1 *)c->v_S.p, (G1 *)c->v_W.p, m);
        c->mark(KZGB200_KC_PAIRING);
        if ((rc = vm_pairing_check(c->stream, c->pairing, (const G1 *)c->v_S.p, 0, (const G1 *)c->v_W.p, 1, d_status, d_status, m))) return rc;   // e(-A, G2) e(pi, [s]G2) == 1
        c->launches += 2;
        c->mark(-1);
        CU(cudaGetLastError());
        if (!st_dev) CU(cudaMemcpyAsync(status + off, d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->marks_collect();
    }
    return KZGB200_OK;
}

int lane_verify_kzg_proof(kzg_lane *c, const uint8_t *commitments48, const uint8_t *z32, const uint8_t *y32, const uint8_t *proofs48, size_t n, int32_t *status) {
    if (n && (!z32 || !y32)) return set_err(KZGB200_ERR_ARGS, "null argument");
    return verify_independent(c, nullptr, commitments48, z32, y32, proofs48, n, status);
}
int lane_verify_blob_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *commitments48, const uint8_t *proofs48, size_t n, int32_t *status) {
    if (n && !blobs) return set_err(KZGB200_ERR_ARGS, "null argument");
    return verify_independent(c, blobs, commitments48, nullptr, nullptr, proofs48, n, status);
}

static void random_scalar_plain(kzg_lane *c, uint32_t *limbs) {   // 248 random bits from the OS entropy source (std::random_device -> /dev/urandom)
    (void)c;
    std::random_device rd;
    for (int i = 0; i < 8; ++i) limbs[i] = rd();
    limbs[7] &= 0x00ffffffu;
    if (!(limbs[0] | limbs[1] | limbs[2] | limbs[3])) limbs[0] = 1;
}

// VerifyBlobKZGProofBatch (verify.go:88-145 + internal/kzg/kzg_verify.go:111-202): one verdict
int lane_verify_blob_kzg_proof_batch(kzg_lane *c, const uint8_t *blobs, const uint8_t *cm48, const uint8_t *pf48, size_t n, int32_t *result) {
    if (!c || !result || (n && (!blobs || !cm48 || !pf48))) return set_err(KZGB200_ERR_ARGS, "null argument");
    if (is_device_ptr(result)) return set_err(KZGB200_ERR_ARGS, "result must be a host pointer");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    *result = KZGB200_OK;
    if (n == 0) return KZGB200_OK;                       // kzg_verify.go:120-122
    int rc;
    // work items of the bucket MSMs: runs of <= 128 points, first the proofs (verdict slot 0), then the commitments (slot 1)
    const uint64_t RLC_ITEM = g_rlc_item > 0 ? (uint64_t)g_rlc_item : KZG_RLC_ITEM_DEFAULT;
    std::vector<uint64_t> item_start, item_end, slot_item_off(3, 0);
    for (int slot = 0; slot < 2; ++slot) {
        for (uint64_t s0 = 0; s0 < n; s0 += RLC_ITEM) { item_start.push_back(slot * n + s0); item_end.push_back(slot * n + std::min<uint64_t>(n, s0 + RLC_ITEM)); }
        slot_item_off[slot + 1] = item_start.size();
    }
    const size_t n_items = item_start.size();
    if ((rc = c->status.ensure(n * sizeof(int32_t)))) return rc;
    if ((rc = c->v_fr.ensure(n * sizeof(Fr)))) return rc;
    if ((rc = c->v_st2.ensure(sizeof(int32_t)))) return rc;
    if ((rc = c->v_aff2.ensure(2 * n * sizeof(G1Aff)))) return rc;       // [proofs | commitments]
    if ((rc = c->zbuf.ensure(n * 32))) return rc;
    if ((rc = c->ybuf.ensure(n * 32))) return rc;
    if ((rc = c->vm_digits.ensure(2 * n * KZG_CELL_TW))) return rc;
    if ((rc = c->vm_scratch.ensure(vm_scratch_bytes(n_items, KZG_CELL_TW, KZG_VM_BUCKETS)))) return rc;
    if ((rc = c->vm_ws.ensure(n_items * KZG_CELL_TW * sizeof(G1)))) return rc;
    if ((rc = c->vm_wsb.ensure(2 * KZG_CELL_TW * sizeof(G1)))) return rc;
    if ((rc = c->v_S.ensure(KZG_VM_SEGS * 2 * sizeof(G1)))) return rc;
    if ((rc = c->v_pa.ensure(sizeof(G1)))) return rc;
    if ((rc = c->v_pb.ensure(sizeof(G1)))) return rc;
    if ((rc = c->v_meta.ensure((2 * n_items + 3) * 8))) return rc;
    if ((rc = c->scalars.ensure(64 * 32))) return rc;
    if ((rc = c->sums.ensure(sizeof(G1)))) return rc;
    G1Aff *d_pf_aff = (G1Aff *)c->v_aff2.p, *d_cm_aff = d_pf_aff + n;
    uint64_t *d_is = (uint64_t *)c->v_meta.p, *d_ie = d_is + n_items, *d_sio = d_ie + n_items;
    CU(cudaMemcpyAsync(d_is, item_start.data(), n_items * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_ie, item_end.data(), n_items * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_sio, slot_item_off.data(), 3 * 8, cudaMemcpyHostToDevice, c->stream));
    // the per-blob front end runs in chunks (bounded staging), all writing into full-size arrays
    for (size_t off = 0; off < n; off += VERIFY_CHUNK) {
        size_t m = std::min(VERIFY_CHUNK, n - off);
        if ((rc = verify_front(c, blobs + off * KZGB200_BYTES_PER_BLOB, cm48 + off * 48, nullptr, nullptr, pf48 + off * 48, m, (int32_t *)c->status.p + off,
                               d_cm_aff + off, d_pf_aff + off, (uint32_t *)c->zbuf.p + off * 8, (uint32_t *)c->ybuf.p + off * 8))) return rc;
        CU(cudaStreamSynchronize(c->stream));
        trace_pt("batch: front synchronised");
    }
    std::vector<int32_t> h_status(n);
    CU(cudaMemcpyAsync(h_status.data(), c->status.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i) if (h_status[i] != KZGB200_OK) { *result = h_status[i]; c->marks_collect(); return KZGB200_OK; }   // first failing element (verify.go:102-119)
    uint32_t seed[8];
    random_scalar_plain(c, seed);
    Fr r_dev; memcpy(r_dev.v, seed, sizeof seed);      // PRF seed for the 126-bit coefficients; n == 1 uses coefficient 1 (kzg_verify.go:125-127)
    c->mark(KZGB200_KC_VERIFY);
    if ((rc = vm_rlc_coeff_digits(c->stream, r_dev, n == 1 ? 1 : 0, (const uint32_t *)c->zbuf.p, (const uint32_t *)c->ybuf.p, (const int32_t *)c->status.p,
                                  (Fr *)c->v_fr.p, (int8_t *)c->vm_digits.p, n))) return rc;
    c->mark(KZGB200_KC_VMSM);
    if ((rc = vm_msm_windows(c->stream, d_pf_aff, (const int8_t *)c->vm_digits.p, KZG_CELL_TW, 0, KZG_CELL_TW, nullptr, KZG_VM_BUCKETS, d_is, d_ie, n_items, d_sio, 2,
                             (G1 *)c->vm_scratch.p, (G1 *)c->vm_ws.p, (G1 *)c->vm_wsb.p))) return rc;
    if ((rc = vm_combine(c->stream, (const G1 *)c->vm_wsb.p, KZG_CELL_TW, 32, 4, KZG_VM_SEGS, (G1 *)c->v_S.p, 2))) return rc;
    c->mark(KZGB200_KC_VERIFY);
    k_rlc_fsum<<<1, 128, 0, c->stream>>>((const Fr *)c->v_fr.p, n, (uint32_t *)c->scalars.p);
    launch_msm_fixed(dim3(1, 1), 32, c->stream, (const uint32_t *)c->scalars.p, c->mono64_tab, 64, 1, 32, nullptr, (G1 *)c->sums.p);   // [sum r_i y_i] G
    k_rlc_prep<<<1, 32, 0, c->stream>>>((const G1 *)c->v_S.p, (const G1 *)c->sums.p, (G1 *)c->v_pa.p, (G1 *)c->v_pb.p);
    c->mark(KZGB200_KC_PAIRING);
    if ((rc = vm_pairing_check(c->stream, c->pairing, (const G1 *)c->v_pa.p, 0, (const G1 *)c->v_pb.p, 1, nullptr, (int32_t *)c->v_st2.p, 1))) return rc;
    c->launches += 9;
    c->mark(-1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(result, c->v_st2.p, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->marks_collect();
    return KZGB200_OK;
}

// VerifyCellKZGProofBatch (api_eip7594.go:163-265 + internal/kzg_multi/kzg_verify.go:16-105)
// cell_indices and batch_offsets (n_batches+1 entries) are read on the host.
int lane_verify_cell_kzg_proof_batch(kzg_lane *c, const uint8_t *commitments48, const uint64_t *cell_indices, const uint8_t *cells,
                                        const uint8_t *proofs48, size_t N, const uint64_t *batch_offsets, size_t nb, int32_t *results) {
    if (!c || (nb && (!batch_offsets || !results)) || (N && (!commitments48 || !cell_indices || !cells || !proofs48)))
        return set_err(KZGB200_ERR_ARGS, "null argument");
    if (N && (is_device_ptr(cell_indices) || is_device_ptr(batch_offsets))) return set_err(KZGB200_ERR_ARGS, "cell_indices/batch_offsets must be host pointers");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    if (nb == 0) return KZGB200_OK;
    const bool res_dev = is_device_ptr(results);
    int rc;
    // ---- optimistic pass: many small verdicts are first checked as ONE random linear combination over all their cells (the
    // large-verdict shape below: coefficients grouped by column, 32 instead of 96 windows per proof, one pairing check instead of
    // nb).  If that single check passes, every verdict is OK (a verdict that should fail survives with probability < 2^-125, the
    // same bound as its own check); anything else -- a false result, a decode or range error anywhere -- says nothing about which
    // verdict is to blame, and the exact per-verdict pass below runs.  Invalid input therefore costs both passes
    // (~1.7 x; KZGB200_OPTIMISTIC=0 or the tunable "optimistic" turns the first pass off).
    static const bool opt_env = [] { const char *e = getenv("KZGB200_OPTIMISTIC"); return !(e && *e == '0'); }();
    if (opt_env && g_optimistic && nb >= 2 && N >= 2 * (size_t)KZG_LARGE_BATCH && N / nb < (size_t)KZG_LARGE_BATCH) {
        bool offsets_ok = batch_offsets[0] == 0 && batch_offsets[nb] == N;
        for (size_t b = 0; b < nb && offsets_ok; ++b) offsets_ok = batch_offsets[b] <= batch_offsets[b + 1];
        if (offsets_ok) {
            const uint64_t one[2] = {0, N};
            int32_t all = -1;
            rc = lane_verify_cell_kzg_proof_batch(c, commitments48, cell_indices, cells, proofs48, N, one, 1, &all);
            if (rc == KZGB200_OK && all == KZGB200_OK) {
                if (res_dev) { CU(cudaMemsetAsync(results, 0, nb * sizeof(int32_t), c->stream)); CU(cudaStreamSynchronize(c->stream)); }
                else memset(results, 0, nb * sizeof(int32_t));
                return KZGB200_OK;
            }
            if (rc != KZGB200_OK) return rc;
            c->timing_reset();
        }
    }
    // ---- gate (tail overlap, kzgb200_api.cu): this call is the SMALL part of a share that was cut in two on one GPU; its kernels wait until the
    // big part's decode has run, so that they fill the GPU while the big part walks its latency-bound tails (Horner, column twiddles, pairing)
    if (c->gate_wait) {
        LaneGate *g = c->gate_wait;
        c->gate_wait = nullptr;                                     // consumed: the optimistic inner call / the exact pass do not wait again
        while (g->state.load(std::memory_order_acquire) == 0) std::this_thread::yield();
        if (g->state.load(std::memory_order_acquire) == 1) {
            CU(cudaStreamWaitEvent(c->stream, g->ev, 0));
            CU(cudaStreamWaitEvent(c->aux_stream, g->ev, 0));
        }
    }
    // ---- device first: the proofs' decode + subgroup check (the longest kernel of the call) needs nothing from
    // the host bookkeeping below, and the cells' H2D rides the copy stream meanwhile ------------------
    const void *d_cells = cells, *d_proofs = nullptr;
    if ((rc = c->v_cst.ensure(std::max<size_t>(N, 1) * 4))) return rc;
    if ((rc = c->v_aff2.ensure(std::max<size_t>(N, 1) * sizeof(G1Aff)))) return rc;
    int32_t *d_cst = (int32_t *)c->v_cst.p;
    c->mark(KZGB200_KC_DECODE);
    if ((rc = stage_in(c, proofs48, N * 48, c->in_small, &d_proofs))) return rc;
    CU(cudaMemsetAsync(d_cst, 0, std::max<size_t>(N, 1) * 4, c->stream));
    if ((rc = vm_g1_check(c->stream, (const uint8_t *)d_proofs, (G1Aff *)c->v_aff2.p, d_cst, N, 1, 1))) return rc;
    if (c->gate_signal && c->gate_signal->state.load(std::memory_order_acquire) == 0) {      // the BIG part: its decode is queued
        CU(cudaEventRecord(c->gate_signal->ev, c->stream));
        c->gate_signal->state.store(1, std::memory_order_release);
    }
    if (N && !is_device_ptr(cells)) {
        if ((rc = c->in_bytes.ensure(N * 2048))) return rc;
        // (every entry point ends with a stream synchronise under the context lock, so the staging buffer is free)
        CU(cudaMemcpyAsync(c->in_bytes.p, cells, N * 2048, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->ev1, c->copy_stream));
        d_cells = c->in_bytes.p;
    }
    // ---- host: commitments for de-duplication -------------------------------------------------
    std::vector<uint8_t> h_cm_copy;
    const uint8_t *h_cm = commitments48;
    if (N && is_device_ptr(commitments48)) {
        h_cm_copy.resize(N * 48);
        CU(cudaMemcpyAsync(h_cm_copy.data(), commitments48, N * 48, cudaMemcpyDeviceToHost, c->copy_stream));
        CU(cudaStreamSynchronize(c->copy_stream));
        h_cm = h_cm_copy.data();
    }
    // (pure host code in cell_plan.hpp; the GPU is already decoding the proofs meanwhile)
    // work items = runs of <= ITEM cells of one verdict (interpolation: one CTA per item; small verdicts' bucket MSM: 96 window-threads per item).
    // A big call is pipe-bound and wants long runs (the bucket reduction costs 16 full additions per item and window); a small call is ONE
    // thread's latency per item -- 128 cells as one item are 128 dependent additions (0.9 ms) -- and wants many short ones: aim at ~30 k tasks.
    uint64_t ITEM = 512;
    while (ITEM > 16 && (N / ITEM) * 96 < 30000) ITEM >>= 1;
    // large verdicts: per (verdict, column) bucket MSM of the 126-bit coefficients.  Default: the SAME signed 4-bit digits as the small
    // verdicts (windows 0..31 of the digit rows), runs of 128 cells -> 32 x N/128 tasks of 128 + 16 additions, one full warp per item.
    // (Round 1 used 16 signed 8-bit windows over runs of 512: 16 x N/512 half-warp tasks of 512 + 256 additions -- 18.3 ms per
    // 524 288 cells, occupancy-bound; tunable "large_window" = 8 brings it back for measurements.)
    const bool l4 = g_large_window != 8;
    const int l_tw = l4 ? 32 : KZG_LARGE_TW, l_nbk = l4 ? KZG_VM_BUCKETS : KZG_LARGE_BUCKETS, l_dbl = l4 ? 4 : 8;
    const uint64_t L_ITEM = g_large_item > 0 ? (uint64_t)g_large_item : (l4 ? 128 : 512);
    CellPlan P;
    if (!plan_cell_batches(h_cm, cell_indices, N, batch_offsets, nb, ITEM, KZG_LARGE_BATCH, KZG_ROW_ITEM, KZGB200_BAD_CELL_INDEX, P, L_ITEM)) {
        cudaStreamSynchronize(c->copy_stream); cudaStreamSynchronize(c->stream);
        return set_err(KZGB200_ERR_ARGS, "batch_offsets not monotone / out of range");
    }
    auto &h_bstatus = P.bstatus; auto &batch_of = P.batch_of; auto &row_cells = P.row_cells; auto &batch_start = P.batch_start;
    auto &batch_row_off = P.batch_row_off; auto &row_off = P.row_off; auto &item_start = P.item_start; auto &item_end = P.item_end;
    auto &batch_item_off = P.batch_item_off; auto &uniq_bytes = P.uniq_bytes; auto &vs_item_start = P.vs_item_start; auto &vs_item_end = P.vs_item_end;
    auto &vs_batch_item_off = P.vs_batch_item_off; auto &l_item_start = P.l_item_start; auto &l_item_end = P.l_item_end; auto &l_slot_item_off = P.l_slot_item_off;
    auto &l_order = P.l_order; auto &large_ids = P.large_ids; auto &r_item_start = P.r_item_start; auto &r_item_end = P.r_item_end;
    auto &r_slot_item_off = P.r_slot_item_off; auto &large_of = P.large_of; auto &row_batch = P.row_batch;
    const size_t U = row_off.size() - 1, n_items = item_start.size();
    const size_t n_large = large_ids.size(), n_vs = vs_item_start.size(), n_li = l_item_start.size(), n_slots = n_large * 128, n_ri = r_item_start.size();
    Fr seed_dev; random_scalar_plain(c, seed_dev.v);   // PRF seed of this call's 126-bit coefficients

    // ---- device buffers ---------------------------------------------------------------------------
    if ((rc = c->in_small2.ensure(std::max<size_t>(U, 1) * 48))) return rc;
    // the unique commitments travel and are decoded on the aux stream, underneath the proofs' decode on the main stream
    if (U) CU(cudaMemcpyAsync(c->in_small2.p, uniq_bytes.data(), U * 48, cudaMemcpyHostToDevice, c->aux_stream));
    // meta arena layout
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = (o + bytes + 255) & ~(size_t)255; return at; };
    size_t o_idx = take(N * 8), o_batch_of = take(N * 4), o_bstart = take(nb * 8);
    size_t o_rowc = take(N * 4), o_rowoff = take((U + 1) * 8), o_browoff = take((nb + 1) * 8), o_is = take(n_items * 8), o_ie = take(n_items * 8);
    size_t o_bio = take((nb + 1) * 8), o_bst = take(nb * 4), o_ust = take(std::max<size_t>(U, 1) * 4);
    size_t o_rowb = take(std::max<size_t>(U, 1) * 4), o_res = take(nb * 4), o_bkey = take(nb * 8);
    // without large verdicts the bucket MSM shares the interpolation's work items
    size_t o_vis = o_is, o_vie = o_ie, o_vbio = o_bio, o_lis = 0, o_lie = 0, o_lsio = 0, o_lord = 0, o_lids = 0, o_ris = 0, o_rie = 0, o_rsio = 0, o_lof = 0;
    if (n_large) {
        o_vis = take(n_vs * 8); o_vie = take(n_vs * 8); o_vbio = take((nb + 1) * 8);
        o_lis = take(n_li * 8); o_lie = take(n_li * 8); o_lsio = take((n_slots + 1) * 8); o_lord = take(l_order.size() * 4); o_lids = take(n_large * 4);
        o_ris = take(n_ri * 8); o_rie = take(n_ri * 8); o_rsio = take((n_large + 1) * 8); o_lof = take(nb * 4);
    }
    if ((rc = c->v_meta.ensure(o))) return rc;
    char *M = (char *)c->v_meta.p;
    // (on the aux stream: a copy from pageable memory waits for the stream it is queued on, and the main stream is busy decoding the proofs)
    auto up = [&](size_t at, const void *src, size_t bytes) { return bytes ? cudaMemcpyAsync(M + at, src, bytes, cudaMemcpyHostToDevice, c->aux_stream) : cudaSuccess; };
    CU(up(o_idx, cell_indices, N * 8)); CU(up(o_batch_of, batch_of.data(), N * 4)); CU(up(o_bstart, batch_start.data(), nb * 8));
    CU(up(o_rowc, row_cells.data(), N * 4));
    CU(up(o_rowoff, row_off.data(), (U + 1) * 8)); CU(up(o_browoff, batch_row_off.data(), (nb + 1) * 8));
    CU(up(o_is, item_start.data(), n_items * 8)); CU(up(o_ie, item_end.data(), n_items * 8)); CU(up(o_bio, batch_item_off.data(), (nb + 1) * 8));
    CU(up(o_bst, h_bstatus.data(), nb * 4)); CU(up(o_rowb, row_batch.data(), U * 4));
    if (n_large) {
        CU(up(o_vis, vs_item_start.data(), n_vs * 8)); CU(up(o_vie, vs_item_end.data(), n_vs * 8)); CU(up(o_vbio, vs_batch_item_off.data(), (nb + 1) * 8));
        CU(up(o_lis, l_item_start.data(), n_li * 8)); CU(up(o_lie, l_item_end.data(), n_li * 8)); CU(up(o_lsio, l_slot_item_off.data(), (n_slots + 1) * 8));
        CU(up(o_lord, l_order.data(), l_order.size() * 4)); CU(up(o_lids, large_ids.data(), n_large * 4));
        CU(up(o_ris, r_item_start.data(), n_ri * 8)); CU(up(o_rie, r_item_end.data(), n_ri * 8)); CU(up(o_rsio, r_slot_item_off.data(), (n_large + 1) * 8));
        CU(up(o_lof, large_of.data(), nb * 4));
    }
    CU(cudaMemsetAsync(M + o_ust, 0, std::max<size_t>(U, 1) * 4, c->aux_stream));
    CU(cudaMemsetAsync(M + o_bkey, 0xff, nb * 8, c->stream));
    if ((rc = c->v_aff1.ensure(std::max<size_t>(U, 1) * sizeof(G1Aff)))) return rc;
    if ((rc = c->v_fr.ensure(std::max<size_t>(N, 1) * sizeof(Fr)))) return rc;
    if ((rc = c->vm_digits.ensure(std::max<size_t>(N, 1) * KZG_CELL_TW))) return rc;
    const size_t n_vm_items = n_large ? n_vs : n_items;
    if ((rc = c->vm_scratch.ensure(std::max<size_t>(std::max(vm_scratch_bytes(n_vm_items, KZG_CELL_TW, KZG_VM_BUCKETS), vm_scratch_bytes(n_li, l_tw, l_nbk)), 256)))) return rc;
    if ((rc = c->vm_ws.ensure(std::max<size_t>(std::max(n_vm_items * KZG_CELL_TW, n_li * (size_t)l_tw), 1) * sizeof(G1)))) return rc;
    if ((rc = c->vm_wsb.ensure(std::max(nb * KZG_CELL_TW, n_slots * (size_t)l_tw) * sizeof(G1)))) return rc;
    if ((rc = c->v_xst.ensure(std::max<size_t>(N, 1) * 4))) return rc;
    if (n_large) {
        if ((rc = c->vm_scratch_r.ensure(std::max<size_t>(vm_scratch_bytes(n_ri, KZG_ROW_TW, KZG_VM_BUCKETS), 256)))) return rc;
        if ((rc = c->vm_ws_r.ensure(std::max<size_t>(n_ri * KZG_ROW_TW, 1) * sizeof(G1)))) return rc;
        if ((rc = c->vm_wsb_r.ensure(n_large * KZG_ROW_TW * sizeof(G1)))) return rc;
        if (!l4 && (rc = c->vm_digits256.ensure(N * KZG_LARGE_TW))) return rc;
        if ((rc = c->vm_colsum.ensure(n_slots * sizeof(G1)))) return rc;
        if ((rc = c->vm_rowdig.ensure(std::max<size_t>(U, 1) * KZG_ROW_TW))) return rc;
        if ((rc = c->vm_commsum.ensure(n_large * sizeof(G1)))) return rc;
    }
    if ((rc = c->v_S.ensure(KZG_VM_SEGS * nb * sizeof(G1)))) return rc;
    if ((rc = c->v_pa.ensure(nb * sizeof(G1)))) return rc;
    if ((rc = c->v_pb.ensure(nb * sizeof(G1)))) return rc;
    if ((rc = c->v_W.ensure(nb * sizeof(G1)))) return rc;
    if ((rc = c->v_partial.ensure(std::max<size_t>(n_items, 1) * 64 * sizeof(Fr)))) return rc;
    if ((rc = c->scalars.ensure(nb * 64 * 32))) return rc;
    if ((rc = c->sums.ensure(nb * sizeof(G1)))) return rc;
    Fr inv64; memcpy(inv64.v, H_FR_INV64, sizeof inv64.v);
    int32_t *d_ust = (int32_t *)(M + o_ust), *d_bst = (int32_t *)(M + o_bst), *d_res = (int32_t *)(M + o_res);
    int32_t *d_xst = (int32_t *)c->v_xst.p;
    // ---- side chain (aux stream, highest priority): everything that does not need the decoded proofs -- the commitments' decode, the
    // coefficients, the cells' interpolation and [sum r_k I_k(s)]G, the commitment weights of large verdicts -- runs BESIDE the proofs'
    // decode of the main stream and fills the issue slots that kernel leaves idle (87 % pipe-active on its own).  Tunable
    // "verify_overlap" = 0 queues the same launches on the main stream instead (round-1 order) for measurements.
    cudaStream_t sa = g_verify_overlap ? c->aux_stream : c->stream;
    if ((rc = vm_g1_check(c->aux_stream, (const uint8_t *)c->in_small2.p, (G1Aff *)c->v_aff1.p, d_ust, U, 1, 1))) return rc;
    CU(cudaMemsetAsync(d_xst, 0, std::max<size_t>(N, 1) * 4, c->aux_stream));
    if (!g_verify_overlap) { CU(cudaEventRecord(c->ev_aux_join, c->aux_stream)); CU(cudaStreamWaitEvent(c->stream, c->ev_aux_join, 0)); }
    if (N) {
        if ((rc = vm_cell_coeff_digits(sa, seed_dev, (const uint32_t *)(M + o_batch_of), (const uint64_t *)(M + o_bstart), (const uint64_t *)(M + o_idx),
                                       c->roots, (Fr *)c->v_fr.p, (int8_t *)c->vm_digits.p, (n_large && !l4) ? (int8_t *)c->vm_digits256.p : nullptr, N))) return rc;
        if (g_verify_overlap) CU(cudaEventRecord(c->ev_fork, sa));                  // the digits are ready: the bucket MSM of the main stream needs nothing else of this chain
        if (d_cells == c->in_bytes.p) CU(cudaStreamWaitEvent(sa, c->ev1, 0));     // the cells have landed
        if (!g_verify_overlap) c->mark(KZGB200_KC_FR);
        k_cell_interp<<<(unsigned)n_items, 256, 0, sa>>>((const uint8_t *)d_cells, (const uint64_t *)(M + o_idx), (const Fr *)c->v_fr.p, (const uint64_t *)(M + o_is),
                                                         (const uint64_t *)(M + o_ie), c->roots, inv64, d_xst, (Fr *)c->v_partial.p);
        c->launches += 3;
    }
    k_cell_interp_reduce<<<(unsigned)nb, 64, 0, sa>>>((const Fr *)c->v_partial.p, (const uint64_t *)(M + o_bio), (uint32_t *)c->scalars.p);
    if (!g_verify_overlap) c->mark(KZGB200_KC_MSM);
    for (size_t b0 = 0; b0 < nb; b0 += 65535) {      // gridDim.y <= 65535
        const size_t bn = std::min<size_t>(65535, nb - b0);
        launch_msm_fixed(dim3(1, (unsigned)bn), 32, sa, (const uint32_t *)c->scalars.p + b0 * 64 * 8, c->mono64_tab, 64, 1, 32, nullptr, (G1 *)c->sums.p + b0);
    }
    CU(cudaGetLastError());
    // sum of w_row C_row of every small verdict: one scalar multiplication per unique commitment, one thread per verdict (latency, hidden here)
    k_cell_row_weights<<<(unsigned)((nb + 31) / 32), 32, 0, sa>>>((const G1Aff *)c->v_aff1.p, (const uint64_t *)(M + o_rowoff), (const uint64_t *)(M + o_browoff),
                                                                  (const uint32_t *)(M + o_rowc), (const Fr *)c->v_fr.p, d_bst, n_large ? (const int32_t *)(M + o_lof) : nullptr,
                                                                  (G1 *)c->v_W.p, nb);
    if (n_large) {
        // sum of w_row C_row over the large verdicts' unique commitments: one more bucket MSM, 40 four-bit windows over short runs
        if ((rc = vm_row_weight_digits(sa, (const Fr *)c->v_fr.p, (const uint64_t *)(M + o_rowoff), (const uint32_t *)(M + o_rowc), (int8_t *)c->vm_rowdig.p, U))) return rc;
        if ((rc = vm_msm_windows(sa, (const G1Aff *)c->v_aff1.p, (const int8_t *)c->vm_rowdig.p, KZG_ROW_TW, 0, KZG_ROW_TW, nullptr, KZG_VM_BUCKETS,
                                 (const uint64_t *)(M + o_ris), (const uint64_t *)(M + o_rie), n_ri, (const uint64_t *)(M + o_rsio), n_large,
                                 (G1 *)c->vm_scratch_r.p, (G1 *)c->vm_ws_r.p, (G1 *)c->vm_wsb_r.p))) return rc;
        if ((rc = vm_combine(sa, (const G1 *)c->vm_wsb_r.p, KZG_ROW_TW, KZG_ROW_TW, 4, 1, (G1 *)c->vm_commsum.p, n_large))) return rc;
    }
    if (g_verify_overlap) { CU(cudaEventRecord(c->ev_aux_join, c->aux_stream)); if (N) CU(cudaStreamWaitEvent(c->stream, c->ev_fork, 0)); }
    c->mark(KZGB200_KC_VMSM);
    // v_S[seg][b]: seg 0 = sum_k r_k pi_k, seg 1 + phi2(seg 2) = sum_k r_k h_k^64 pi_k   (kzg_verify.go:32,73-83)
    if ((rc = vm_msm_windows(c->stream, (const G1Aff *)c->v_aff2.p, (const int8_t *)c->vm_digits.p, KZG_CELL_TW, 0, KZG_CELL_TW, nullptr, KZG_VM_BUCKETS,
                             (const uint64_t *)(M + o_vis), (const uint64_t *)(M + o_vie), n_vm_items, (const uint64_t *)(M + o_vbio), nb,
                             (G1 *)c->vm_scratch.p, (G1 *)c->vm_ws.p, (G1 *)c->vm_wsb.p))) return rc;
    if ((rc = vm_combine(c->stream, (const G1 *)c->vm_wsb.p, KZG_CELL_TW, 32, 4, KZG_VM_SEGS, (G1 *)c->v_S.p, nb))) return rc;
    if (n_large) {
        // large verdicts: per (verdict, column) bucket MSM of the coefficients, then the column twiddles
        if ((rc = vm_msm_windows(c->stream, (const G1Aff *)c->v_aff2.p, l4 ? (const int8_t *)c->vm_digits.p : (const int8_t *)c->vm_digits256.p, l4 ? KZG_CELL_TW : KZG_LARGE_TW, 0, l_tw,
                                 (const uint32_t *)(M + o_lord), l_nbk, (const uint64_t *)(M + o_lis), (const uint64_t *)(M + o_lie), n_li, (const uint64_t *)(M + o_lsio), n_slots,
                                 (G1 *)c->vm_scratch.p, (G1 *)c->vm_ws.p, (G1 *)c->vm_wsb.p))) return rc;
        if ((rc = vm_combine(c->stream, (const G1 *)c->vm_wsb.p, l_tw, l_tw, l_dbl, 1, (G1 *)c->vm_colsum.p, n_slots))) return rc;
        if ((rc = vm_cell_columns_large(c->stream, (const G1 *)c->vm_colsum.p, (const uint32_t *)(M + o_lids), n_large, c->glv_digits, (G1 *)c->v_S.p, nb))) return rc;
        c->launches += 11;
    }
    if (g_verify_overlap) CU(cudaStreamWaitEvent(c->stream, c->ev_aux_join, 0));      // the rest of the side chain: statuses, [sum r I(s)]G, commitment weights
    c->mark(KZGB200_KC_VERIFY);
    unsigned long long *d_bkey = (unsigned long long *)(M + o_bkey);
    if (N) k_merge_status<<<(unsigned)((N + 127) / 128), 128, 0, c->stream>>>(d_cst, (const uint32_t *)(M + o_batch_of), d_bkey, N, 1);      // the proofs' decode errors (stage 1)
    if (N) k_merge_status<<<(unsigned)((N + 127) / 128), 128, 0, c->stream>>>(d_xst, (const uint32_t *)(M + o_batch_of), d_bkey, N, 1);      // the cells' NON_CANONICAL_SCALAR (stage 2)
    if (U) k_merge_status<<<(unsigned)((U + 127) / 128), 128, 0, c->stream>>>(d_ust, (const uint32_t *)(M + o_rowb), d_bkey, U, 0);
    k_status_finish<<<(unsigned)((nb + 127) / 128), 128, 0, c->stream>>>(d_bkey, d_bst, nb);
    k_cell_prep<<<(unsigned)((nb + 31) / 32), 32, 0, c->stream>>>((const G1 *)c->v_S.p, (const G1 *)c->sums.p, (const G1 *)c->v_W.p,
                                                                   d_bst, n_large ? (const int32_t *)(M + o_lof) : nullptr, (const G1 *)c->vm_commsum.p, (G1 *)c->v_pa.p, (G1 *)c->v_pb.p, nb);
    c->mark(KZGB200_KC_PAIRING);
    if ((rc = vm_pairing_check(c->stream, c->pairing, (const G1 *)c->v_pa.p, 2, (const G1 *)c->v_pb.p, 0, d_bst, d_res, nb))) return rc;
    c->launches += 10;   // + bucket reduce and item reduce inside vm_msm_windows
    c->mark(-1);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(results, d_res, nb * 4, res_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->marks_collect();
    return KZGB200_OK;
}

// unit-test hook (host only, no GPU): the plan of plan_cell_batches as JSON text (thread-local buffer), or NULL on malformed offsets
const char *kzgb200_dbg_plan_cell_batches_json_l(const uint8_t *commitments48, const uint64_t *cell_indices, size_t N, const uint64_t *batch_offsets, size_t nb,
                                                 uint64_t item, uint64_t large, uint64_t row_item, uint64_t l_item);
const char *kzgb200_dbg_plan_cell_batches_json(const uint8_t *commitments48, const uint64_t *cell_indices, size_t N, const uint64_t *batch_offsets, size_t nb,
                                               uint64_t item, uint64_t large, uint64_t row_item) {
    return kzgb200_dbg_plan_cell_batches_json_l(commitments48, cell_indices, N, batch_offsets, nb, item, large, row_item, 0);
}
// the same with the run length of the large verdicts' column items given separately (0 = item)
const char *kzgb200_dbg_plan_cell_batches_json_l(const uint8_t *commitments48, const uint64_t *cell_indices, size_t N, const uint64_t *batch_offsets, size_t nb,
                                                 uint64_t item, uint64_t large, uint64_t row_item, uint64_t l_item) {
    static thread_local std::string out;
    CellPlan P;
    if (!plan_cell_batches(commitments48, cell_indices, N, batch_offsets, nb, item, large, row_item, KZGB200_BAD_CELL_INDEX, P, l_item)) return nullptr;
    out.clear();
    auto arr = [&](const char *name, auto &v, bool last = false) {
        out += "\""; out += name; out += "\": [";
        for (size_t i = 0; i < v.size(); ++i) { if (i) out += ","; out += std::to_string((long long)v[i]); }
        out += last ? "]" : "],";
    };
    out += "{";
    arr("bstatus", P.bstatus); arr("batch_start", P.batch_start); arr("batch_row_off", P.batch_row_off); arr("batch_item_off", P.batch_item_off);
    arr("large_of", P.large_of); arr("batch_of", P.batch_of); arr("uniq_bytes", P.uniq_bytes); arr("row_off", P.row_off); arr("row_cells", P.row_cells);
    arr("row_batch", P.row_batch); arr("item_start", P.item_start); arr("item_end", P.item_end); arr("vs_item_start", P.vs_item_start);
    arr("vs_item_end", P.vs_item_end); arr("vs_batch_item_off", P.vs_batch_item_off); arr("large_ids", P.large_ids); arr("l_order", P.l_order);
    arr("l_item_start", P.l_item_start); arr("l_item_end", P.l_item_end); arr("l_slot_item_off", P.l_slot_item_off); arr("r_item_start", P.r_item_start);
    arr("r_item_end", P.r_item_end); arr("r_slot_item_off", P.r_slot_item_off, true);
    out += "}";
    return out.c_str();
}

int kzgb200_dbg_g1_mul(const uint8_t *p48, const uint8_t *s32, uint8_t *out48, int n) {
    uint8_t *dp, *ds, *dout;
    CU(cudaMalloc(&dp, n * 48)); CU(cudaMalloc(&ds, n * 32)); CU(cudaMalloc(&dout, n * 48));
    CU(cudaMemcpy(dp, p48, n * 48, cudaMemcpyHostToDevice)); CU(cudaMemcpy(ds, s32, n * 32, cudaMemcpyHostToDevice));
    k_dbg_g1_mul<<<(n + 31) / 32, 32>>>(dp, ds, dout, n);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out48, dout, n * 48, cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(ds); cudaFree(dout);
    return 0;
}
int lane_dbg_pairing(kzg_lane *c, const uint8_t *a48, const int *qa, const uint8_t *b48, const int *qb, int *out, int n) {
    uint8_t *da, *db; int32_t *dout; G1 *dA, *dB;
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(&da, n * 48)); CU(cudaMalloc(&db, n * 48)); CU(cudaMalloc(&dout, n * 4)); CU(cudaMalloc(&dA, n * sizeof(G1))); CU(cudaMalloc(&dB, n * sizeof(G1)));
    CU(cudaMemcpy(da, a48, n * 48, cudaMemcpyHostToDevice)); CU(cudaMemcpy(db, b48, n * 48, cudaMemcpyHostToDevice));
    k_dbg_decode_pair<<<(n + 31) / 32, 32>>>(da, db, dA, dB, n);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    int rc = 0;
    for (int i = 0; i < n && !rc; ++i) rc = vm_pairing_check(nullptr, c->pairing, dA + i, qa[i], dB + i, qb[i], nullptr, dout + i, 1);   // selectors are per launch
    if (rc) return rc;
    std::vector<int32_t> st(n);
    CU(cudaMemcpy(st.data(), dout, n * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) out[i] = st[i] == KZGB200_OK ? 1 : 0;
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dA); cudaFree(dB);
    return 0;
}
int lane_dbg_dump_pairing(kzg_lane *c, uint32_t *out96) {
    uint32_t *d;
    CU(cudaSetDevice(c->device));
    CU(cudaMalloc(&d, 96 * 4));
    k_dbg_dump_pairing<<<1, 1>>>(c->pairing, d);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out96, d, 96 * 4, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

}  // extern "C"
