// Shared host-side definitions of libkzgb200.so: error plumbing, the LANE object, staging helpers.
//
// A lane (kzg_lane) is one execution slot on one GPU: its own stream set, event pool and grow-only scratch arena,
// plus pointers to the read-only tables of its GPU (owned by the first lane created there, shared by its clones).
// The public context (kzgb200_ctx, kzgb200_api.cu) is a set of lanes per GPU; it hands every call -- or every
// device's share of a call -- to a free lane, so a context is usable from many host threads at once and over
// several GPUs.  The lane_* functions below are the single-GPU implementations of the entry points of
// include/kzgb200.h; they are not locked: the pool guarantees a lane runs one call at a time.
// Included by every translation unit: kzgb200.cu (setup + proving paths), kzgb200_verify.cu (host side of the
// verifiers, built with -Xptxas -O1 -- see build.py for why), kzgb200_vmsm.cu (verifier throughput kernels, launch
// wrappers declared at the bottom of this file), kzgb200_setup.cu (trusted-setup ingestion, codec validation).
#pragma once
#include "../../include/kzgb200.h"
#include "../../include/kzgb200_debug.h"
#include "msm.cuh"
#include "fk20.cuh"
#include "kzg4844.cuh"
#include "recover.cuh"
#include "pairing.cuh"
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <thread>

using namespace kzg;

std::string &kzgb200_err_slot();   // thread-local last-error text (defined in kzgb200.cu)
static inline int set_err(int code, const char *what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    kzgb200_err_slot() = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return set_err(KZGB200_ERR_CUDA, #call, e_); } while (0)

static const int N_BLOB = 4096;

namespace kzg {
// decompress + (optional) subgroup check of n compressed points; first error per status slot wins.
// out may be null (validation only: prove.go:56-60 discards the point).
template <class M_ = MulCall, int MINB = 4> static __global__ void __launch_bounds__(64, MINB) k_g1_check(const uint8_t *__restrict__ in48, G1Aff *__restrict__ out, int32_t *__restrict__ status, size_t n, int per_status, int subgroup) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Aff a;
    int32_t st = g1_decompress(a, in48 + i * 48);
    if (st == ST_OK && subgroup && !g1_in_subgroup<M_>(&a, FP_BETA2)) st = ST_NOT_IN_SUBGROUP;
    if (st != ST_OK) atomicCAS(&status[i / per_status], (int32_t)ST_OK, st);
    if (out) out[i] = a;
}

}  // namespace kzg

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return set_err(KZGB200_ERR_CUDA, "cudaMalloc(scratch)", e);
        cap = bytes;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// device temporaries of context creation: freed on every exit path (they leaked on the error returns)
struct InitTemps {
    std::vector<void *> ptrs;
    template <class T> void track(T *p) { ptrs.push_back((void *)p); }
    ~InitTemps() { for (void *p : ptrs) if (p) cudaFree(p); }
};

// A gate between two lanes of one GPU (kzgb200_api.cu: tail overlap of the cell verifier): the first lane records `ev` on its main stream once its
// long decode kernel is queued, the second lane makes its kernel streams wait for it.  state: 0 = nothing yet, 1 = ev recorded, 2 = released without one.
struct LaneGate { cudaEvent_t ev = nullptr; std::atomic<int> state{0}; };

#define KZG_G1FFT_MAX_SPLIT 8
struct kzg_lane {
    LaneGate *gate_signal = nullptr, *gate_wait = nullptr;      // set by the public layer around ONE call, consumed by lane_verify_cell_kzg_proof_batch
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_piece[8] = {nullptr};
    cudaStream_t fft_streams[KZG_G1FFT_MAX_SPLIT - 1] = {nullptr};   // sub-batches of the staged G1 FFT (launch_fk20_proofs)
    cudaEvent_t ev_fork = nullptr, ev_join[KZG_G1FFT_MAX_SPLIT - 1] = {nullptr};
    cudaStream_t aux_stream = nullptr;  // small independent kernels that would otherwise queue behind a long one (second decode of the verifiers)
    cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
    size_t g1fft_split = 8;             // measured: 1 -> 40.9 ms, 2 -> 34.6, 4 -> 33.5, 8 -> 33.1 (KZGB200_G1FFT_SPLIT overrides)
    size_t g1_chain4_max = 64;          // ... then up to this many the 4 x 4 x 4 x 2 form (KZGB200_G1_CHAIN4_MAX overrides).  Measured, g1fft class in ms for
                                        // two-level / 4x4x4x2 / staged: 25-32 blobs 9.2 / 10.3 / 16.2, 33-64: 15.5 / 12.6 / 16.2, 65-96: 22.0 / 18.5 / 16.2, 128: 27.3 / 20.1 / 16.3
    size_t g1_two_level_max = 32;       // ... up to this many the two-level 16 x 8 form, beyond it the staged radix-2 form (KZGB200_G1_TWO_LEVEL_MAX overrides).
                                        // Measured (profiles/r02_g1_midbatch_sweep.json), two-level / staged: 25-32 blobs 9.2 / 16.2 ms, 33-64: 15.5 / 16.2, 96: 22.0 / 16.2
    size_t g1_dense_max = 24;           // batches up to this many blobs take the dense one-level G1 transform (KZGB200_G1_DENSE_MAX overrides)
    bool owns_tables = true;            // false for clones: setup pointers below belong to the GPU's first lane
    // setup
    G1Aff *g1_monomial = nullptr;      // natural order
    G1Aff *g1_lagrange_brp = nullptr;  // bit-reversed order (api.go:131)
    std::vector<uint8_t> g2_bytes;
    MsmTable commit_tab{};
    MsmTable fk20_tab{};
    Fr *roots = nullptr;               // w_8192^t, Montgomery
    int8_t *glv_digits = nullptr;      // [128][2][KZG_GLV_DIGITS]
    Fr *pow7 = nullptr, *ipow7 = nullptr;   // 7^k, 7^-k (erasure_code.go:58 coset generator)
    PairingConsts *pairing = nullptr;       // Frobenius constants + line tables of G2, [s]G2, [s^64]G2
    MsmTable mono64_tab{};                  // digit table of monomial G1[0..63] (kzg_multi/srs.go:143-149)
    // scratch
    DevBuf msm_partial;
    DevBuf in_bytes, scalars, status, sums, out_bytes, coeffs, cells, proofs_xyzz, fft_work, in_small, in_small2, zbuf, ybuf;
    DevBuf rec_a, rec_b, rec_meta, rec_zev, rec_czinv;
    DevBuf v_aff1, v_aff2, v_fr, v_meta, v_S, v_W, v_partial, v_in2, v_in3, v_st2;
    DevBuf vm_digits, vm_digits256, vm_colsum, vm_rowdig, vm_commsum, vm_scratch, vm_ws, vm_wsb, v_pa, v_pb, v_cst, v_st3, v_pst, ev_cex, ev_total, ev_index;
    DevBuf v_xst, vm_scratch_r, vm_ws_r, vm_wsb_r;      // cell verifier: the cells' own status slots; scratch of the row-weight MSM (runs beside the column MSM)
    double init_ms = 0, last_device_ms = 0;
    uint64_t launches = 0;
    // per-kernel-class device timing of the last call (CUDA events on `stream`)
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> mark_cls;
    size_t n_marks = 0;
    double class_ms[KZGB200_N_KERNEL_CLASSES] = {0};
    void marks_reset() { n_marks = 0; mark_cls.clear(); }
    // the segment that starts here belongs to kernel class `cls` (-1 = end marker)
    int mark(int cls) {
        if (n_marks == ev_pool.size()) {
            cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return 1;
            ev_pool.push_back(e);
        }
        if (cudaEventRecord(ev_pool[n_marks], stream) != cudaSuccess) return 1;
        mark_cls.push_back(cls); ++n_marks;
        return 0;
    }
    void marks_collect() {   // stream must be synchronised
        for (size_t i = 0; i + 1 < n_marks; ++i) {
            float ms = 0;
            if (mark_cls[i] >= 0 && cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]) == cudaSuccess) {
                class_ms[mark_cls[i]] += ms; last_device_ms += ms;
            }
        }
        marks_reset();
    }
    void timing_reset() { last_device_ms = 0; for (double &x : class_ms) x = 0; marks_reset(); }
};

static inline bool is_device_ptr(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// stage a (possibly host) input buffer on the device
static inline int stage_in(kzg_lane *c, const void *user, size_t bytes, DevBuf &buf, const void **dev) {
    if (bytes == 0) { *dev = buf.p; return 0; }          // empty input: user may be null
    if (is_device_ptr(user)) { *dev = user; return 0; }
    int rc = buf.ensure(bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(buf.p, user, bytes, cudaMemcpyHostToDevice, c->stream));
    *dev = buf.p;
    return 0;
}

// ---- per-lane implementations of the C ABI (kzgb200.cu, kzgb200_verify.cu, kzgb200_setup.cu) ----
extern "C" {
int lane_ctx_new(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2_monomial, size_t n_g2, const kzgb200_opts *opts, int device, kzg_lane **out);
int lane_clone(kzg_lane *first, kzg_lane **out);      // another lane on the same GPU sharing `first`'s tables
void lane_ctx_free(kzg_lane *c);
void lane_quiesce(kzg_lane *c);       // after a failed call: wait for everything the call left in flight on the lane's streams, drop its event marks
int lane_blob_to_kzg_commitment(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out48, int32_t *status);
int lane_compute_blob_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *commitments48, size_t n, uint8_t *out48, int32_t *status);
int lane_compute_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *z32, size_t n, uint8_t *out_proof48, uint8_t *out_y32, int32_t *status);
int lane_verify_kzg_proof(kzg_lane *c, const uint8_t *commitments48, const uint8_t *z32, const uint8_t *y32, const uint8_t *proofs48, size_t n, int32_t *status);
int lane_verify_blob_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *commitments48, const uint8_t *proofs48, size_t n, int32_t *status);
int lane_verify_blob_kzg_proof_batch(kzg_lane *c, const uint8_t *blobs, const uint8_t *commitments48, const uint8_t *proofs48, size_t n, int32_t *result);
int lane_compute_cells(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out_cells, int32_t *status);
int lane_compute_cells_and_kzg_proofs(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out_cells, uint8_t *out_proofs, int32_t *status);
int lane_recover_cells_and_kzg_proofs(kzg_lane *c, const uint64_t *cell_ids, const uint64_t *counts, const uint8_t *cells, size_t n,
                                      uint8_t *out_cells, uint8_t *out_proofs, int32_t *status);
int lane_verify_cell_kzg_proof_batch(kzg_lane *c, const uint8_t *commitments48, const uint64_t *cell_indices, const uint8_t *cells, const uint8_t *proofs48,
                                     size_t n_cells, const uint64_t *batch_offsets, size_t n_batches, int32_t *results);
int lane_check_g1_points(kzg_lane *c, const uint8_t *points48, size_t n, int32_t *status);
int lane_check_scalars(kzg_lane *c, const uint8_t *scalars32, size_t n_items, size_t scalars_per_item, int32_t *status);
int lane_dbg_pairing(kzg_lane *c, const uint8_t *a48, const int *qa, const uint8_t *b48, const int *qb, int *out, int n);
int lane_dbg_dump_pairing(kzg_lane *c, uint32_t *out96);
int lane_dbg_vmsm(kzg_lane *c, const uint8_t *p48, const uint32_t *s, int n, uint8_t *out48);
}

// ---- launch wrappers of kzgb200_vmsm.cu (verification throughput kernels, full optimisation) ----
int vm_g1_check(cudaStream_t st, const uint8_t *in48, G1Aff *out, int32_t *status, size_t n, int per_status, int subgroup);
int vm_cell_coeff_digits(cudaStream_t st, const Fr &seed, const uint32_t *batch_of, const uint64_t *batch_start, const uint64_t *cell_idx,
                         const Fr *roots, Fr *rpow, int8_t *digits, int8_t *digits256, size_t n);
size_t vm_scratch_bytes(size_t n_items, int nw, int nbuckets);
int vm_msm_windows(cudaStream_t st, const G1Aff *points, const int8_t *digits, int TW, int w_lo, int nw, const uint32_t *order, int nbuckets,
                   const uint64_t *item_start, const uint64_t *item_end, size_t n_items, const uint64_t *batch_item_off, size_t nb,
                   G1 *scratch, G1 *WS, G1 *WSb);
// out[seg * nb + b] = Horner over windows [nw seg, nw seg + nw) of WSb[b], dbl doublings between windows
int vm_combine(cudaStream_t st, const G1 *WSb, int TW, int nw, int dbl, int n_segs, G1 *out, size_t nb);
int vm_row_weight_digits(cudaStream_t st, const Fr *rpow, const uint64_t *row_off, const uint32_t *row_cells, int8_t *digits, size_t n_rows);
int vm_cell_columns_large(cudaStream_t st, const G1 *colsum, const uint32_t *large_ids, size_t n_large, const int8_t *glv_tw_digits, G1 *comb, size_t nb);
int vm_rlc_coeff_digits(cudaStream_t st, const Fr &seed, int unit_coeff, const uint32_t *z, const uint32_t *y, const int32_t *status,
                        Fr *fy, int8_t *digits, size_t n);
// result[i] = pre_status[i] if that is an error (pre_status may be null or alias result), else OK / VERIFY_FAILED for
// e(A_i, Q[qa]) e(B_i, Q[qb]) == 1, Q = {G2, [s]G2, [s^64]G2}
int vm_pairing_check(cudaStream_t st, const PairingConsts *pc, const G1 *A, int qa, const G1 *B, int qb, const int32_t *pre_status, int32_t *result, size_t n);
// Open / EvaluateLagrangePolynomial for m blobs on stream st (k_eval_products, k_fr_inv_batch, k_eval_finish), using
// scratch slots [slot, slot + m) of the scratch sized by vm_eval_scratch(total)
int vm_eval_scratch(kzg_lane *c, size_t total);
int vm_eval_quotient(kzg_lane *c, cudaStream_t st, size_t slot, const uint8_t *d_blobs, const uint32_t *z_limbs, int32_t *d_status, uint32_t *quotient,
                     uint8_t *y_out, uint32_t *y_limbs, size_t m);
