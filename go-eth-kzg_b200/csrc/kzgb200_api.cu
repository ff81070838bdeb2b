// libkzgb200.so -- the public context (include/kzgb200.h): a pool of execution lanes per GPU over one or several GPUs.
//
// The reference's Context is immutable and is called concurrently from many goroutines, one blob per call
// (api.go:17-28, verify.go:159-166); its batch verifier fans out over goroutines itself (verify.go:152-169).
// This layer gives the CUDA engine the same two properties:
//   * concurrency: every GPU has `lanes` lanes (own streams + scratch, shared read-only tables); a call takes a free
//     lane, small calls rotate over the GPUs, so N host threads calling single-blob methods overlap on the device(s);
//   * multi-GPU in ONE process: a batched call on host buffers is cut into contiguous ranges of
//     independent units (shard_plan.hpp), one host thread per GPU copies and computes its range and writes its slice
//     of the caller's output buffers; no collective, no peer traffic.  An RLC verdict is computed as one sub-verdict
//     per GPU, each with its own random coefficients, merged on the host (first error in index order, else AND).
// The per-GPU work is the lane_* functions of kzgb200.cu / kzgb200_verify.cu / kzgb200_setup.cu.
#include "ctx.cuh"
#include "shard_plan.hpp"
#include <atomic>
#include <condition_variable>
#include <memory>
#include <thread>

namespace {

struct DevSlot {
    int device = 0;
    std::vector<kzg_lane *> lanes;
    std::vector<kzg_lane *> free_list;
    std::mutex mu;
    std::condition_variable cv;
    kzg_lane *acquire() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return !free_list.empty(); });
        kzg_lane *l = free_list.back(); free_list.pop_back();
        return l;
    }
    void release(kzg_lane *l) {
        { std::lock_guard<std::mutex> lk(mu); free_list.push_back(l); }
        cv.notify_one();
    }
};

// minimum units per GPU before a host-buffer call is cut (below it the hand-off and the partial waves cost more than
// the second GPU brings; KZGB200_SPLIT_MIN scales all of them, 0 disables splitting)
struct SplitMin { size_t commit = 64, cells = 16, verify_blob = 64, verify_cells = 4096; };

}  // namespace

struct kzgb200_ctx {
    std::vector<std::unique_ptr<DevSlot>> devs;
    std::atomic<unsigned> rr{0};
    int lanes_per_dev = 2;
    SplitMin split;
    double init_ms = 0;
    std::mutex rec_mu;
    double last_device_ms = 0, class_ms[KZGB200_N_KERNEL_CLASSES] = {0};
};

namespace {

struct LaneRef {
    DevSlot *d; kzg_lane *l;
    explicit LaneRef(DevSlot *d_) : d(d_), l(d_->acquire()) {}
    ~LaneRef() { d->release(l); }
    LaneRef(const LaneRef &) = delete;
};

// CUDA device that owns p, or -1 for host memory
int ptr_device(const void *p) {
    if (!p) return -1;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}

// index in ctx->devs of the GPU that owns one of the buffers, -1 if all are host memory, -2 if a buffer lives on a
// GPU this context does not use
int owning_dev(kzgb200_ctx *ctx, std::initializer_list<const void *> ptrs) {
    for (const void *p : ptrs) {
        int d = ptr_device(p);
        if (d < 0) continue;
        for (size_t i = 0; i < ctx->devs.size(); ++i) if (ctx->devs[i]->device == d) return (int)i;
        return -2;
    }
    return -1;
}

void record(kzgb200_ctx *ctx, const std::vector<kzg_lane *> &used) {
    std::lock_guard<std::mutex> lk(ctx->rec_mu);
    ctx->last_device_ms = 0;
    for (double &x : ctx->class_ms) x = 0;
    for (kzg_lane *l : used) {
        ctx->last_device_ms = std::max(ctx->last_device_ms, l->last_device_ms);
        for (int k = 0; k < KZGB200_N_KERNEL_CLASSES; ++k) ctx->class_ms[k] = std::max(ctx->class_ms[k], l->class_ms[k]);
    }
}

// Runs f(lane, part index, range) for every range: one range -> on the calling thread, on the GPU `pinned_dev` (>= 0) or the
// next GPU in rotation; several ranges -> range i on GPU i, range 0 on the calling thread, the others on threads of their own.
// Returns the first non-zero code in range order (its text becomes this thread's kzgb200_last_error).
template <class F> int run_ranges(kzgb200_ctx *ctx, const std::vector<ShardRange> &ranges, int pinned_dev, F f) {
    const size_t n_dev = ctx->devs.size();
    if (ranges.empty()) return KZGB200_OK;
    if (ranges.size() == 1) {
        size_t d = pinned_dev >= 0 ? (size_t)pinned_dev : (n_dev == 1 ? 0 : ctx->rr.fetch_add(1) % n_dev);
        LaneRef ref(ctx->devs[d].get());
        int rc = f(ref.l, (size_t)0, ranges[0]);
        if (rc) { std::string keep = kzgb200_err_slot(); lane_quiesce(ref.l); kzgb200_err_slot() = keep; }
        record(ctx, {ref.l});
        return rc;
    }
    std::vector<int> rcs(ranges.size(), 0);
    std::vector<std::string> errs(ranges.size());
    std::vector<kzg_lane *> used(ranges.size(), nullptr);
    auto body = [&](size_t i) {
        LaneRef ref(ctx->devs[i % n_dev].get());
        used[i] = ref.l;
        rcs[i] = f(ref.l, i, ranges[i]);
        if (rcs[i]) { errs[i] = kzgb200_err_slot(); lane_quiesce(ref.l); }
    };
    std::vector<std::thread> th;
    for (size_t i = 1; i < ranges.size(); ++i) th.emplace_back(body, i);
    body(0);
    for (auto &t : th) t.join();
    record(ctx, used);
    for (size_t i = 0; i < ranges.size(); ++i) if (rcs[i]) { kzgb200_err_slot() = errs[i]; return rcs[i]; }
    return KZGB200_OK;
}

// ranges of a call over n independent units: one range if any buffer is device memory (it runs where the buffers live)
int plan(kzgb200_ctx *ctx, size_t n, size_t min_per_dev, std::initializer_list<const void *> ptrs, std::vector<ShardRange> &ranges, int &pinned) {
    pinned = n ? owning_dev(ctx, ptrs) : -1;
    if (pinned == -2) return set_err(KZGB200_ERR_ARGS, "a device buffer lives on a GPU that is not part of this context");
    if (pinned >= 0 || ctx->devs.size() == 1 || min_per_dev == 0) ranges.assign(1, ShardRange{0, n});
    else ranges = shard_units(n, ctx->devs.size(), min_per_dev);
    if (ranges.empty()) ranges.assign(1, ShardRange{0, 0});
    return 0;
}

}  // namespace

extern "C" {

int kzgb200_ctx_new(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2_monomial, size_t n_g2,
                    const kzgb200_opts *opts, kzgb200_ctx **out) {
    if (!g1_monomial || !g1_lagrange || !g2_monomial || !out) return set_err(KZGB200_ERR_ARGS, "null argument");
    *out = nullptr;
    auto t0 = std::chrono::steady_clock::now();
    int visible = 0;
    CU(cudaGetDeviceCount(&visible));
    std::vector<int> ordinals;
    const int nd = opts ? opts->n_devices : 0;
    if (nd == 0) ordinals.push_back(opts ? opts->device : 0);
    else if (nd < 0) for (int i = 0; i < visible; ++i) ordinals.push_back(i);
    else {
        if (!opts->devices) return set_err(KZGB200_ERR_ARGS, "n_devices > 0 needs a devices array");
        ordinals.assign(opts->devices, opts->devices + nd);
    }
    if (ordinals.empty()) return set_err(KZGB200_ERR_CUDA, "no CUDA device");
    for (size_t i = 0; i < ordinals.size(); ++i) {
        if (ordinals[i] < 0 || ordinals[i] >= visible) return set_err(KZGB200_ERR_CUDA, "no such CUDA device");
    }
    // (an ordinal listed twice gets two replicas and two lane pools on that GPU: useless in production, but it lets a
    // one-GPU box run the multi-device path -- tests/test_gpu_multidev.py)
    std::unique_ptr<kzgb200_ctx> ctx(new kzgb200_ctx());
    int lanes = opts && opts->lanes ? opts->lanes : 0;
    if (!lanes) { const char *e = getenv("KZGB200_LANES"); lanes = e ? atoi(e) : 2; }
    ctx->lanes_per_dev = std::min(std::max(lanes, 1), 16);
    if (const char *e = getenv("KZGB200_SPLIT_MIN")) {
        const double f = atof(e);
        ctx->split.commit = (size_t)(ctx->split.commit * f); ctx->split.cells = (size_t)(ctx->split.cells * f);
        ctx->split.verify_blob = (size_t)(ctx->split.verify_blob * f); ctx->split.verify_cells = (size_t)(ctx->split.verify_cells * f);
    }
    // every GPU builds its own replica of the tables, concurrently
    const size_t n_dev = ordinals.size();
    std::vector<int> rcs(n_dev, 0);
    std::vector<std::string> errs(n_dev);
    for (size_t i = 0; i < n_dev; ++i) { ctx->devs.emplace_back(new DevSlot()); ctx->devs[i]->device = ordinals[i]; }
    auto build = [&](size_t i) {
        DevSlot *d = ctx->devs[i].get();
        kzg_lane *first = nullptr;
        int rc = lane_ctx_new(g1_monomial, g1_lagrange, g2_monomial, n_g2, opts, d->device, &first);
        if (!rc) {
            d->lanes.push_back(first);
            for (int k = 1; k < ctx->lanes_per_dev && !rc; ++k) {
                kzg_lane *c = nullptr;
                rc = lane_clone(first, &c);
                if (!rc) d->lanes.push_back(c);
            }
        }
        if (rc) errs[i] = kzgb200_err_slot();
        rcs[i] = rc;
        d->free_list = d->lanes;
    };
    std::vector<std::thread> th;
    for (size_t i = 1; i < n_dev; ++i) th.emplace_back(build, i);
    build(0);
    for (auto &t : th) t.join();
    for (size_t i = 0; i < n_dev; ++i) if (rcs[i]) {
        kzgb200_err_slot() = errs[i];
        int rc = rcs[i];
        kzgb200_ctx_free(ctx.release());
        return rc;
    }
    ctx->init_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = ctx.release();
    return KZGB200_OK;
}

void kzgb200_ctx_free(kzgb200_ctx *ctx) {
    if (!ctx) return;
    for (auto &d : ctx->devs) {
        // clones first: the first lane of a GPU owns the tables
        for (size_t k = d->lanes.size(); k-- > 0;) lane_ctx_free(d->lanes[k]);
    }
    delete ctx;
}

int kzgb200_get_info(kzgb200_ctx *ctx, kzgb200_info *o) {
    if (!ctx || !o) return set_err(KZGB200_ERR_ARGS, "null argument");
    kzg_lane *c = ctx->devs[0]->lanes[0];
    memset(o, 0, sizeof *o);
    o->device = c->device; o->sm_count = c->sm_count;
    o->commit_window = c->commit_tab.c; o->commit_windows_per_scalar = c->commit_tab.W;
    o->commit_table_bytes = c->commit_tab.bytes();
    o->fk20_window = c->fk20_tab.c; o->fk20_windows_per_scalar = c->fk20_tab.W;
    o->fk20_table_bytes = c->fk20_tab.bytes();
    o->init_ms = ctx->init_ms;
    for (auto &d : ctx->devs) for (kzg_lane *l : d->lanes) o->kernel_launches += l->launches;
    o->n_devices = (int)ctx->devs.size(); o->lanes_per_device = ctx->lanes_per_dev;
    return KZGB200_OK;
}
double kzgb200_last_device_ms(kzgb200_ctx *ctx) {
    if (!ctx) return 0.0;
    std::lock_guard<std::mutex> lk(ctx->rec_mu);
    return ctx->last_device_ms;
}
int kzgb200_last_kernel_ms(kzgb200_ctx *ctx, double *out) {
    if (!ctx || !out) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::lock_guard<std::mutex> lk(ctx->rec_mu);
    for (int i = 0; i < KZGB200_N_KERNEL_CLASSES; ++i) out[i] = ctx->class_ms[i];
    return KZGB200_OK;
}

// ---- per-blob independent calls: contiguous blob ranges ----------------------------------------------------
int kzgb200_blob_to_kzg_commitment(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out48, int32_t *status) {
    if (!ctx || (n && (!blobs || !out48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.commit, {blobs, out48, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_blob_to_kzg_commitment(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, r.hi - r.lo, out48 + r.lo * 48, status + r.lo);
    });
}

int kzgb200_compute_blob_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48, size_t n, uint8_t *out48, int32_t *status) {
    if (!ctx || (n && (!blobs || !commitments48 || !out48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.commit, {blobs, commitments48, out48, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_compute_blob_kzg_proof(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, commitments48 + r.lo * 48, r.hi - r.lo, out48 + r.lo * 48, status + r.lo);
    });
}

int kzgb200_compute_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *z32, size_t n, uint8_t *out_proof48, uint8_t *out_y32, int32_t *status) {
    if (!ctx || (n && (!blobs || !z32 || !out_proof48 || !out_y32 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.commit, {blobs, z32, out_proof48, out_y32, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_compute_kzg_proof(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, z32 + r.lo * 32, r.hi - r.lo, out_proof48 + r.lo * 48, out_y32 + r.lo * 32, status + r.lo);
    });
}

int kzgb200_verify_kzg_proof(kzgb200_ctx *ctx, const uint8_t *commitments48, const uint8_t *z32, const uint8_t *y32, const uint8_t *proofs48, size_t n, int32_t *status) {
    if (!ctx || (n && (!commitments48 || !z32 || !y32 || !proofs48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, 4 * ctx->split.verify_blob, {commitments48, z32, y32, proofs48, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_verify_kzg_proof(l, commitments48 + r.lo * 48, z32 + r.lo * 32, y32 + r.lo * 32, proofs48 + r.lo * 48, r.hi - r.lo, status + r.lo);
    });
}

int kzgb200_verify_blob_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48, const uint8_t *proofs48, size_t n, int32_t *status) {
    if (!ctx || (n && (!blobs || !commitments48 || !proofs48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.verify_blob, {blobs, commitments48, proofs48, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_verify_blob_kzg_proof(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, commitments48 + r.lo * 48, proofs48 + r.lo * 48, r.hi - r.lo, status + r.lo);
    });
}

// one logical verdict = one sub-verdict per GPU, each with its own random coefficients
int kzgb200_verify_blob_kzg_proof_batch(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48, const uint8_t *proofs48, size_t n, int32_t *result) {
    if (!ctx || !result || (n && (!blobs || !commitments48 || !proofs48))) return set_err(KZGB200_ERR_ARGS, "null argument");
    if (ptr_device(result) >= 0) return set_err(KZGB200_ERR_ARGS, "result must be a host pointer");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.verify_blob, {blobs, commitments48, proofs48}, rg, pin))) return rc;
    std::vector<int32_t> sub(rg.size(), KZGB200_OK);
    rc = run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t i, ShardRange r) {
        return lane_verify_blob_kzg_proof_batch(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, commitments48 + r.lo * 48, proofs48 + r.lo * 48, r.hi - r.lo, &sub[i]);
    });
    if (rc) return rc;
    *result = merge_sub_verdicts(sub.data(), sub.size());
    return KZGB200_OK;
}

int kzgb200_compute_cells(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out_cells, int32_t *status) {
    if (!ctx || (n && (!blobs || !out_cells || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, 4 * ctx->split.cells, {blobs, out_cells, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_compute_cells(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, r.hi - r.lo, out_cells + r.lo * 262144, status + r.lo);
    });
}

int kzgb200_compute_cells_and_kzg_proofs(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) {
    if (!ctx || (n && (!blobs || !out_cells || !out_proofs || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.cells, {blobs, out_cells, out_proofs, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_compute_cells_and_kzg_proofs(l, blobs + r.lo * KZGB200_BYTES_PER_BLOB, r.hi - r.lo, out_cells + r.lo * 262144, out_proofs + r.lo * 6144, status + r.lo);
    });
}

int kzgb200_recover_cells_and_kzg_proofs(kzgb200_ctx *ctx, const uint64_t *cell_ids, const uint64_t *counts, const uint8_t *cells, size_t n,
                                         uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) {
    if (!ctx || (n && (!counts || !out_cells || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    if (n && (ptr_device(cell_ids) >= 0 || ptr_device(counts) >= 0)) return set_err(KZGB200_ERR_ARGS, "cell_ids/counts must be host pointers");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, ctx->split.cells, {cells, out_cells, out_proofs, status}, rg, pin))) return rc;
    // blob ranges -> offsets into the flat id / cell arrays
    std::vector<uint64_t> first_cell(rg.size(), 0);
    if (rg.size() > 1) {
        uint64_t acc = 0; size_t k = 0;
        for (size_t b = 0; b < n && k < rg.size(); ++b) { if (b == rg[k].lo) first_cell[k++] = acc; acc += counts[b]; }
    }
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t i, ShardRange r) {
        const uint64_t c0 = first_cell[i];
        return lane_recover_cells_and_kzg_proofs(l, cell_ids ? cell_ids + c0 : nullptr, counts + r.lo, cells ? cells + c0 * 2048 : nullptr, r.hi - r.lo,
                                                 out_cells + r.lo * 262144, out_proofs ? out_proofs + r.lo * 6144 : nullptr, status + r.lo);
    });
}

int kzgb200_verify_cell_kzg_proof_batch(kzgb200_ctx *ctx, const uint8_t *commitments48, const uint64_t *cell_indices, const uint8_t *cells,
                                        const uint8_t *proofs48, size_t N, const uint64_t *batch_offsets, size_t nb, int32_t *results) {
    if (!ctx || (nb && (!batch_offsets || !results)) || (N && (!commitments48 || !cell_indices || !cells || !proofs48)))
        return set_err(KZGB200_ERR_ARGS, "null argument");
    if (N && (ptr_device(cell_indices) >= 0 || ptr_device(batch_offsets) >= 0)) return set_err(KZGB200_ERR_ARGS, "cell_indices/batch_offsets must be host pointers");
    if (nb == 0) { if (N) return set_err(KZGB200_ERR_ARGS, "cells without a verdict"); return KZGB200_OK; }
    if (batch_offsets[0] != 0 || batch_offsets[nb] != N) return set_err(KZGB200_ERR_ARGS, "batch_offsets must cover exactly [0, n_cells)");
    for (size_t b = 0; b < nb; ++b) if (batch_offsets[b + 1] < batch_offsets[b]) return set_err(KZGB200_ERR_ARGS, "batch_offsets not monotone");
    int pin = N ? owning_dev(ctx, {commitments48, cells, proofs48, results}) : owning_dev(ctx, {results});
    if (pin == -2) return set_err(KZGB200_ERR_ARGS, "a device buffer lives on a GPU that is not part of this context");
    const size_t n_dev = ctx->devs.size();
    const bool may_split = pin < 0 && n_dev > 1 && ctx->split.verify_cells > 0;
    const bool one = nb == 1;          // one verdict: the units below are CELLS (sub-verdicts merged at the end), else VERDICTS
    // ---- shares: one per GPU (host buffers, several GPUs), else the whole call on the GPU that owns the buffers / the next in rotation
    std::vector<ShardRange> rg;
    if (may_split && one && N >= 2 * ctx->split.verify_cells) rg = shard_units(N, n_dev, ctx->split.verify_cells);
    else if (may_split && !one) rg = shard_verdicts(batch_offsets, nb, n_dev, ctx->split.verify_cells);
    else rg.assign(1, ShardRange{0, one ? N : nb});
    const size_t solo_dev = pin >= 0 ? (size_t)pin : (n_dev == 1 ? 0 : ctx->rr.fetch_add(1) % n_dev);
    // ---- tail overlap: a share of >= 128 Ki cells is cut once more, 7/8 + 1/8, onto two lanes of the SAME GPU.  The verdict of a big share ends in
    // ~8 ms of dependent chains (Horner over the windows, column twiddles, one pairing) during which the GPU idles; the small part's kernels are
    // gated behind the big part's decode (LaneGate) and fill exactly that time.  MEASURED on 524 288 cells: 47.35 ms against 47.2 ms on one lane --
    // the small part's decode saturates the multiply pipe of every SM, and the big part's one-warp chains, sharing schedulers with it, stretch by
    // as much as the overlap saves (vmsm class 11.2 -> 15.5 ms).  Kept behind the tunable "tail_split" (default OFF) with its test.
    struct Part { size_t dev; ShardRange r; LaneGate *sig, *wait; };
    std::vector<Part> parts;
    std::vector<std::unique_ptr<LaneGate>> gates;
    auto cells_of = [&](const ShardRange &r) { return one ? (uint64_t)(r.hi - r.lo) : batch_offsets[r.hi] - batch_offsets[r.lo]; };
    for (size_t i = 0; i < rg.size(); ++i) {
        const size_t dev = rg.size() > 1 ? i % n_dev : solo_dev;
        const ShardRange r = rg[i];
        size_t cut = r.hi;
        if (kzg::g_tail_split && ctx->lanes_per_dev >= 2 && cells_of(r) >= (uint64_t)kzg::g_tail_min_cells) {
            if (one) cut = r.lo + (r.hi - r.lo) / 8 * 7;
            else {
                const uint64_t target = batch_offsets[r.lo] + cells_of(r) / 8 * 7;
                cut = (size_t)(std::lower_bound(batch_offsets + r.lo, batch_offsets + r.hi, target) - batch_offsets);
            }
        }
        if (cut > r.lo && cut < r.hi) {
            gates.emplace_back(new LaneGate());
            cudaSetDevice(ctx->devs[dev]->device);
            if (cudaEventCreateWithFlags(&gates.back()->ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); gates.back()->state = 2; }
            parts.push_back({dev, {r.lo, cut}, gates.back().get(), nullptr});
            parts.push_back({dev, {cut, r.hi}, nullptr, gates.back().get()});
        } else parts.push_back({dev, r, nullptr, nullptr});
    }
    std::vector<int32_t> sub(parts.size(), KZGB200_OK);
    std::vector<int> rcs(parts.size(), 0);
    std::vector<std::string> errs(parts.size());
    std::vector<kzg_lane *> used(parts.size(), nullptr);
    auto body = [&](size_t i) {
        const Part &p = parts[i];
        LaneRef ref(ctx->devs[p.dev].get());
        kzg_lane *l = ref.l;
        used[i] = l;
        l->gate_signal = p.sig; l->gate_wait = p.wait;
        if (one) {
            const uint64_t off[2] = {0, p.r.hi - p.r.lo};
            rcs[i] = lane_verify_cell_kzg_proof_batch(l, commitments48 + p.r.lo * 48, cell_indices + p.r.lo, cells + p.r.lo * 2048, proofs48 + p.r.lo * 48,
                                                      p.r.hi - p.r.lo, off, 1, rg.size() == 1 && parts.size() == 1 ? results : &sub[i]);
        } else {
            const uint64_t c0 = batch_offsets[p.r.lo], c1 = batch_offsets[p.r.hi];
            std::vector<uint64_t> off(p.r.hi - p.r.lo + 1);
            for (size_t b = p.r.lo; b <= p.r.hi; ++b) off[b - p.r.lo] = batch_offsets[b] - c0;
            rcs[i] = lane_verify_cell_kzg_proof_batch(l, commitments48 + c0 * 48, cell_indices + c0, cells + c0 * 2048, proofs48 + c0 * 48, c1 - c0, off.data(),
                                                      p.r.hi - p.r.lo, results + p.r.lo);
        }
        l->gate_signal = nullptr; l->gate_wait = nullptr;
        if (p.sig && p.sig->state.load() == 0) p.sig->state.store(2);          // the big part ended without queueing its decode: release the small one
        if (rcs[i]) { errs[i] = kzgb200_err_slot(); lane_quiesce(l); }
    };
    {
        std::vector<std::thread> th;
        for (size_t i = 1; i < parts.size(); ++i) th.emplace_back(body, i);
        body(0);
        for (auto &t : th) t.join();
    }
    for (auto &g : gates) if (g->ev) cudaEventDestroy(g->ev);
    record(ctx, used);
    for (size_t i = 0; i < parts.size(); ++i) if (rcs[i]) { kzgb200_err_slot() = errs[i]; return rcs[i]; }
    if (one && parts.size() > 1) {
        // sub-verdicts of ONE verdict (each with its own random coefficients), merged in index order
        const int32_t merged = merge_sub_verdicts(sub.data(), sub.size());
        if (ptr_device(results) >= 0) { if (cudaMemcpy(results, &merged, sizeof merged, cudaMemcpyHostToDevice) != cudaSuccess) return set_err(KZGB200_ERR_CUDA, "cudaMemcpy(result)"); }
        else results[0] = merged;
    }
    return KZGB200_OK;
}

int kzgb200_check_g1_points(kzgb200_ctx *ctx, const uint8_t *points48, size_t n, int32_t *status) {
    if (!ctx || (n && (!points48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n, 64 * ctx->split.verify_cells, {points48, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) { return lane_check_g1_points(l, points48 + r.lo * 48, r.hi - r.lo, status + r.lo); });
}
int kzgb200_check_scalars(kzgb200_ctx *ctx, const uint8_t *scalars32, size_t n_items, size_t scalars_per_item, int32_t *status) {
    if (!ctx || !scalars_per_item || (n_items && (!scalars32 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    std::vector<ShardRange> rg; int pin, rc;
    if ((rc = plan(ctx, n_items, 0, {scalars32, status}, rg, pin))) return rc;
    return run_ranges(ctx, rg, pin, [&](kzg_lane *l, size_t, ShardRange r) {
        return lane_check_scalars(l, scalars32 + r.lo * scalars_per_item * 32, r.hi - r.lo, scalars_per_item, status + r.lo);
    });
}

// ---- debug hooks that need a context (include/kzgb200_debug.h): first lane of the first GPU --------------------
int kzgb200_dbg_pairing(kzgb200_ctx *ctx, const uint8_t *a48, const int *qa, const uint8_t *b48, const int *qb, int *out, int n) {
    if (!ctx) return set_err(KZGB200_ERR_ARGS, "null argument");
    LaneRef ref(ctx->devs[0].get());
    return lane_dbg_pairing(ref.l, a48, qa, b48, qb, out, n);
}
int kzgb200_dbg_dump_pairing(kzgb200_ctx *ctx, uint32_t *out96) {
    if (!ctx) return set_err(KZGB200_ERR_ARGS, "null argument");
    LaneRef ref(ctx->devs[0].get());
    return lane_dbg_dump_pairing(ref.l, out96);
}
int kzgb200_dbg_vmsm(kzgb200_ctx *ctx, const uint8_t *p48, const uint32_t *s, int n, uint8_t *out48) {
    if (!ctx) return set_err(KZGB200_ERR_ARGS, "null argument");
    LaneRef ref(ctx->devs[0].get());
    return lane_dbg_vmsm(ref.l, p48, s, n, out48);
}

// the shard planner as JSON text (host only, no GPU): {"units": [[lo,hi],...], "verdicts": [[lo,hi],...], "merged": m}
const char *kzgb200_dbg_shard_plan_json(size_t n_units, size_t n_dev, size_t min_per_dev, const uint64_t *batch_offsets, size_t n_batches,
                                        size_t min_cells_per_dev, const int32_t *sub_verdicts, size_t n_sub) {
    static thread_local std::string out;
    auto ranges = [&](const std::vector<ShardRange> &r) {
        out += "[";
        for (size_t i = 0; i < r.size(); ++i) { if (i) out += ","; out += "[" + std::to_string(r[i].lo) + "," + std::to_string(r[i].hi) + "]"; }
        out += "]";
    };
    out = "{\"units\": ";
    ranges(shard_units(n_units, n_dev, min_per_dev));
    out += ", \"verdicts\": ";
    ranges(batch_offsets ? shard_verdicts(batch_offsets, n_batches, n_dev, min_cells_per_dev) : std::vector<ShardRange>());
    out += ", \"merged\": " + std::to_string(sub_verdicts ? merge_sub_verdicts(sub_verdicts, n_sub) : 0) + "}";
    return out.c_str();
}

}  // extern "C"
