// libkzgb200.so -- host side of the C ABI (include/kzgb200.h): context creation and the
// commit / prove / cells / recovery entry points.  The verification entry points live in
// kzgb200_verify.cu.
//
// One context = one GPU.  All per-call scratch lives in a grow-only device arena owned by the
// context; inputs are processed in chunks so the arena stays bounded regardless of batch size.
#define KZG_MSM_ALL_VARIANTS 1
#include "ctx.cuh"
#include "g1fft.cuh"
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

std::string &kzgb200_err_slot() { static thread_local std::string s; return s; }

// ===========================================================================================
// kernels that belong to the API layer
// ===========================================================================================
namespace kzg {

// trusted-setup ingestion: decompress n points; dst index optionally bit-reversed (api.go:131)
static __global__ void k_setup_decompress(const uint8_t *__restrict__ in, G1Aff *__restrict__ out, int n, int log_n_brp, int32_t *__restrict__ bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Aff a;
    int32_t st = g1_decompress(a, in + (size_t)i * 48);
    if (st != ST_OK) atomicMax(bad, st);
    int dst = log_n_brp ? (int)(__brev((unsigned)i) >> (32 - log_n_brp)) : i;
    out[dst] = a;
}

// blob bytes -> plain little-endian scalar limbs + canonical check (serialization.go:134-146)
// one thread per scalar; status[blob] = NON_CANONICAL if any scalar >= r
static __global__ void k_blob_to_scalars(const uint8_t *__restrict__ blobs, uint32_t *__restrict__ scalars, int32_t *__restrict__ status, size_t n_scalars, int per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_scalars) return;
    uint32_t l[8];
    load_be32(l, blobs + i * 32);
    if (!fr_is_canonical(l)) atomicMax(&status[i / per_item], (int32_t)ST_NON_CANONICAL_SCALAR);
    uint4 *o = reinterpret_cast<uint4 *>(scalars + i * 8);
    o[0] = make_uint4(l[0], l[1], l[2], l[3]);
    o[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

// XYZZ sums -> 48-byte compressed points; items with a non-OK status get zero bytes.
// Each thread normalises FIN_BATCH consecutive points with one shared inversion (Montgomery's
// trick over the ZZZ coordinates), so the Fermat inversion is amortised 8x.
#define KZG_FIN_BATCH 8
static __global__ void k_finalize_g1(const G1 *__restrict__ in, uint8_t *__restrict__ out48, const int32_t *__restrict__ status, size_t n, int per_status) {
    size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * KZG_FIN_BATCH;
    if (base >= n) return;
    int cnt = (int)min((size_t)KZG_FIN_BATCH, n - base);
    Fp pre[KZG_FIN_BATCH];
    bool live[KZG_FIN_BATCH];
    Fp run = Fp::one();
    for (int k = 0; k < cnt; ++k) {
        size_t i = base + k;
        Fp zzz = in[i].ZZZ;
        live[k] = !(status && status[i / per_status] != ST_OK) && !in[i].ZZ.is_zero();
        pre[k] = run;
        if (live[k]) run = fp_mul_ni(run, zzz);
    }
    Fp inv = fp_inv(run);
    for (int k = cnt - 1; k >= 0; --k) {
        size_t i = base + k;
        uint8_t *o = out48 + i * 48;
        if (status && status[i / per_status] != ST_OK) {
            uint32_t *w = reinterpret_cast<uint32_t *>(o);
            for (int q = 0; q < 12; ++q) w[q] = 0;
            continue;
        }
        G1Aff a;
        if (!live[k]) { a.x = Fp::zero(); a.y = Fp::zero(); }        // point at infinity
        else {
            G1 p = in[i];
            Fp i3 = fp_mul_ni(inv, pre[k]);                            // 1 / ZZZ_i
            inv = fp_mul_ni(inv, p.ZZZ);
            Fp i2 = fp_mul_ni(fp_sqr_ni(i3), fp_sqr_ni(p.ZZ));         // 1 / ZZ_i
            a.x = fp_mul_ni(p.X, i2);
            a.y = fp_mul_ni(p.Y, i3);
        }
        g1_compress(o, a);
    }
}

// status[i] = status[i] if that is an error, else later[i]
static __global__ void k_status_then(int32_t *__restrict__ status, const int32_t *__restrict__ later, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == ST_OK) status[i] = later[i];
}
// zero the y outputs of failed items
static __global__ void k_zero_failed(uint8_t *__restrict__ out, const int32_t *__restrict__ status, size_t n, int bytes_per_item) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] == ST_OK) return;
    for (int k = 0; k < bytes_per_item; ++k) out[i * bytes_per_item + k] = 0;
}

// ---- debug / unit-test kernels (include/kzgb200_debug.h) ----------------------------------
static __global__ void k_dbg_fp_mul(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp x, y;
    for (int k = 0; k < 12; ++k) { x.v[k] = a[i * 12 + k]; y.v[k] = b[i * 12 + k]; }
    x = Fp::to_mont(x); y = Fp::to_mont(y);
    Fp r = op == 0 ? Fp::mul(x, y) : op == 1 ? Fp::add(x, y) : op == 2 ? Fp::sub(x, y) : op == 3 ? fp_inv(x) : op == 5 ? Fp::mul_add_mul(x, y, Fp::add(x, y), Fp::sub(x, y)) : Fp::sqr(x);
    r = Fp::from_mont(r);
    for (int k = 0; k < 12; ++k) out[i * 12 + k] = r.v[k];
}
static __global__ void k_dbg_fr_mul(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x, y;
    for (int k = 0; k < 8; ++k) { x.v[k] = a[i * 8 + k]; y.v[k] = b[i * 8 + k]; }
    x = Fr::to_mont(x); y = Fr::to_mont(y);
    Fr r = op == 0 ? Fr::mul(x, y) : op == 1 ? Fr::add(x, y) : op == 2 ? Fr::sub(x, y) : Fr::sqr(x);
    r = Fr::from_mont(r);
    for (int k = 0; k < 8; ++k) out[i * 8 + k] = r.v[k];
}
// out = a (op) b on compressed points: op 0 add (XYZZ+XYZZ), 1 mixed add, 2 double
static __global__ void k_dbg_g1_op(const uint8_t *a48, const uint8_t *b48, uint8_t *out48, int n, int op) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Aff a, b;
    g1_decompress(a, a48 + i * 48);
    g1_decompress(b, b48 + i * 48);
    G1 A = G1::from_affine(a);
    if (op == 0) { G1 B = G1::from_affine(b); G1 B2 = g1_dbl(B); g1_add(B2, B); /* 3b, non-trivial ZZ */
                   G1 nb = B; nb.neg_inplace(); g1_add(B2, nb); g1_add(B2, nb); g1_add(A, B2); }
    else if (op == 1) g1_add_affine(A, b);
    else A = g1_dbl(A);
    g1_compress(out48 + i * 48, g1_to_affine(A));
}

// dependency-free integer multiply-add throughput probes: 8 independent chains per thread.
// MODE 0: mad.lo.u32   1: mad.hi.u32   2: mad.wide.u32 (32x32+64 -> 64, one IMAD.WIDE)
template <int MODE> static __global__ void k_imad_peak(uint32_t *out, int iters, uint32_t seed) {
    uint32_t m = blockIdx.x * 2654435761u + 12345u + seed;
    if (MODE < 2) {
        uint32_t a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = seed * (2 * k + 3) + threadIdx.x;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 7]));
                    else asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[k]) : "r"(m), "r"(a[(k + 1) & 7]));
                }
            }
        }
        uint32_t x = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) x ^= a[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    } else {
        unsigned long long a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = (unsigned long long)seed * (2 * k + 3) + threadIdx.x;
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t lo = (uint32_t)a[k];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(lo), "r"(m));
                }
            }
        }
        unsigned long long x = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) x ^= a[k];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(x ^ (x >> 32));
    }
}

// isolated chains of the library's own field operations: the practical ceiling of a kernel made of nothing but Fp products
// (the pipe probe of profiles/r02_pipe_probe.md, kept in the library so bench.py measures it live on the box it runs on)
template <bool SQR> static __global__ void __launch_bounds__(128) k_fp_chain(uint32_t *out, int iters, uint32_t seed) {
    Fp x, y;
#pragma unroll
    for (int k = 0; k < 12; ++k) { x.v[k] = seed * (2 * k + 3) + threadIdx.x + blockIdx.x * 977u; y.v[k] = seed * (2 * k + 5) ^ (threadIdx.x * 2654435761u); }
    x.v[11] &= 0x0fffffffu; y.v[11] &= 0x0fffffffu;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) x = SQR ? Fp::sqr(x) : Fp::mul(x, y);
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 12; ++k) r ^= x.v[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

}  // namespace kzg

static int build_table(kzg_lane *c, const G1Aff *pts, int npts, int window, MsmTable &tab) {
    tab.plan(npts, window);
    size_t n_bases = (size_t)npts * tab.W;
    G1 *bases = nullptr; G1Aff *bases_aff = nullptr;
    InitTemps temps;
    CU(cudaMalloc(&bases, n_bases * sizeof(G1))); temps.track(bases);
    CU(cudaMalloc(&bases_aff, n_bases * sizeof(G1Aff))); temps.track(bases_aff);
    CU(cudaMalloc(&tab.entries, tab.bytes()));
    k_table_bases<<<(npts + 63) / 64, 64, 0, c->stream>>>(pts, tab, bases);
    k_to_affine<<<(unsigned)((n_bases + 127) / 128), 128, 0, c->stream>>>(bases, bases_aff, n_bases);
    size_t threads = (size_t)npts * (tab.row_entries / KZG_TABLE_CHUNK);
    k_table_fill<<<(unsigned)((threads + 63) / 64), 64, 0, c->stream>>>(bases_aff, tab);
    c->launches += 3;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" {

const char *kzgb200_last_error(void) { return kzgb200_err_slot().c_str(); }

void *kzgb200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
// Pinned memory whose pages are interleaved over all NUMA nodes of the host (mmap + mbind(MPOL_INTERLEAVE) +
// cudaHostRegister): a buffer that is read by the GPUs of BOTH sockets at once (a multi-GPU context sharding one host
// batch) then draws on every memory controller instead of the first-touch node's.  Falls back to kzgb200_host_alloc.
static std::mutex g_mapped_mu;
static std::unordered_map<void *, size_t> g_mapped;
void *kzgb200_host_alloc_interleaved(size_t bytes) {
    const size_t len = (bytes + 4095) & ~(size_t)4095;
    void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return kzgb200_host_alloc(bytes);
    unsigned long mask[16];
    for (unsigned long &m : mask) m = ~0ul;                      // every possible node; the kernel intersects with the allowed set
    long rc = syscall(SYS_mbind, p, len, 3 /* MPOL_INTERLEAVE */, mask, (unsigned long)(sizeof mask * 8), 0u);
    if (rc != 0) {                                               // e.g. no NUMA support / seccomp: try the online nodes only
        unsigned long one[1] = {0};
        FILE *f = fopen("/sys/devices/system/node/online", "r");
        int lo = 0, hi = 0;
        if (f) { if (fscanf(f, "%d-%d", &lo, &hi) < 2) hi = lo; fclose(f); }
        for (int n = lo; n <= hi && n < 64; ++n) one[0] |= 1ul << n;
        rc = hi > lo ? syscall(SYS_mbind, p, len, 3, one, 65ul, 0u) : -1;
    }
    memset(p, 0, len);                                           // fault the pages in under the policy
    if (cudaHostRegister(p, len, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError(); munmap(p, len);
        return kzgb200_host_alloc(bytes);
    }
    std::lock_guard<std::mutex> lk(g_mapped_mu);
    g_mapped[p] = len;
    return p;
}
void kzgb200_host_free(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_mapped_mu);
        auto it = g_mapped.find(p);
        if (it != g_mapped.end()) {
            cudaHostUnregister(p); munmap(p, it->second);
            g_mapped.erase(it);
            return;
        }
    }
    cudaFreeHost(p);
}

// aggregate host -> device bandwidth of n GPUs copying AT THE SAME TIME, each its own slice of one pinned host buffer
// (plain or NUMA-interleaved): the ceiling of every end-to-end number of a multi-GPU context
int kzgb200_dbg_h2d_bandwidth(const int *devices, int n, size_t bytes_per_dev, int interleaved, double *gbps) {
    if (!devices || n <= 0 || !gbps) return set_err(KZGB200_ERR_ARGS, "null argument");
    uint8_t *h = (uint8_t *)(interleaved ? kzgb200_host_alloc_interleaved(bytes_per_dev * n) : kzgb200_host_alloc(bytes_per_dev * n));
    if (!h) return set_err(KZGB200_ERR_CUDA, "host allocation failed");
    std::vector<void *> d(n, nullptr);
    std::vector<cudaStream_t> st(n);
    for (int i = 0; i < n; ++i) { CU(cudaSetDevice(devices[i])); CU(cudaMalloc(&d[i], bytes_per_dev)); CU(cudaStreamCreate(&st[i])); }
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        for (int i = 0; i < n; ++i) { cudaSetDevice(devices[i]); cudaStreamSynchronize(st[i]); }
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n; ++i) { cudaSetDevice(devices[i]); cudaMemcpyAsync(d[i], h + (size_t)i * bytes_per_dev, bytes_per_dev, cudaMemcpyHostToDevice, st[i]); }
        for (int i = 0; i < n; ++i) { cudaSetDevice(devices[i]); cudaStreamSynchronize(st[i]); }
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rep) best = std::max(best, (double)bytes_per_dev * n / s / 1e9);
    }
    for (int i = 0; i < n; ++i) { cudaSetDevice(devices[i]); cudaFree(d[i]); cudaStreamDestroy(st[i]); }
    kzgb200_host_free(h);
    *gbps = best;
    return 0;
}

// unit-test hook (host only): source index and twiddle exponent of term j of output o at `level` of the multi-level G1 transforms
// (g1fft.cuh; form 0: 16 x 8, four levels; form 1: 4 x 4 x 4 x 2, eight levels)
int kzgb200_dbg_g1_level_term(int form, int level, int o, int j, int *src, int *e) {
    if (!src || !e || form < 0 || form > 1 || level < 0 || level >= (form ? 8 : 4) || o < 0 || o > 127 || j < 0 || j > 15) return set_err(KZGB200_ERR_ARGS, "bad argument");
    if (form) kzg::g1lvl8_term(level, o, j, *src, *e); else kzg::g1lvl_term(level, o, j, *src, *e);
    return 0;
}

// experiments: run-time tunables of the proving paths.  "msm_variant": the k_msm_fixed variant (msm.cuh; -1 = default /
// KZGB200_MSM_VARIANT); "fk20_lanes": lanes per 64-point FK20 group for full batches (4, 8, 16; 0 = default);
// "vmsm_policy": field-product policy of the verifiers' bucket accumulation (vmsm.cuh)
int kzgb200_dbg_set_tunable(const char *name, int v) {
    if (!name) return set_err(KZGB200_ERR_ARGS, "null argument");
    if (!strcmp(name, "msm_variant")) {
        if (v < -1 || v > 15) return set_err(KZGB200_ERR_ARGS, "variant out of range");
        kzg::g_msm_variant_override = v;
        return 0;
    }
    if (!strcmp(name, "fk20_lanes")) {
        if (v != 0 && v != 4 && v != 8 && v != 16) return set_err(KZGB200_ERR_ARGS, "fk20_lanes must be 0, 4, 8 or 16");
        kzg::g_fk20_lanes_override = v;
        return 0;
    }
    if (!strcmp(name, "g1fft_dual")) {
        if (v != 0 && v != 1) return set_err(KZGB200_ERR_ARGS, "g1fft_dual must be 0 or 1");
        kzg::g_g1fft_dual = v;
        return 0;
    }
    if (!strcmp(name, "decode_dual")) {
        if (v != 0 && v != 1) return set_err(KZGB200_ERR_ARGS, "decode_dual must be 0 or 1");
        kzg::g_decode_dual = v;
        return 0;
    }
    if (!strcmp(name, "large_window")) {
        if (v != 4 && v != 8) return set_err(KZGB200_ERR_ARGS, "large_window must be 4 or 8");
        kzg::g_large_window = v;
        return 0;
    }
    if (!strcmp(name, "large_item")) {
        if (v < 0 || v > 65536) return set_err(KZGB200_ERR_ARGS, "large_item out of range");
        kzg::g_large_item = v;
        return 0;
    }
    if (!strcmp(name, "optimistic")) {
        if (v != 0 && v != 1) return set_err(KZGB200_ERR_ARGS, "optimistic must be 0 or 1");
        kzg::g_optimistic = v;
        return 0;
    }
    if (!strcmp(name, "verify_overlap")) {
        if (v != 0 && v != 1) return set_err(KZGB200_ERR_ARGS, "verify_overlap must be 0 or 1");
        kzg::g_verify_overlap = v;
        return 0;
    }
    if (!strcmp(name, "decode_minb")) {
        if (v != 4 && v != 6 && v != 8) return set_err(KZGB200_ERR_ARGS, "decode_minb must be 4, 6 or 8");
        kzg::g_decode_minb = v;
        return 0;
    }
    if (!strcmp(name, "g1_chain4_max")) {
        if (v < -1 || v > 4096) return set_err(KZGB200_ERR_ARGS, "g1_chain4_max out of range");
        kzg::g_g1_chain4_max = v;
        return 0;
    }
    if (!strcmp(name, "g1_two_level_max")) {
        if (v < -1 || v > 1024) return set_err(KZGB200_ERR_ARGS, "g1_two_level_max out of range");
        kzg::g_g1_two_level_max = v;
        return 0;
    }
    if (!strcmp(name, "proof_pieces")) {
        if (v < 2 || v > 4) return set_err(KZGB200_ERR_ARGS, "proof_pieces out of range");     // KZG_H2D_PIECES
        kzg::g_proof_pieces = v;
        return 0;
    }
    if (!strcmp(name, "tail_split")) {
        if (v != 0 && v != 1) return set_err(KZGB200_ERR_ARGS, "tail_split must be 0 or 1");
        kzg::g_tail_split = v;
        return 0;
    }
    if (!strcmp(name, "tail_min_cells")) {
        if (v < 16) return set_err(KZGB200_ERR_ARGS, "tail_min_cells out of range");
        kzg::g_tail_min_cells = v;
        return 0;
    }
    if (!strcmp(name, "rlc_item")) {
        if (v < 0 || v > 4096) return set_err(KZGB200_ERR_ARGS, "rlc_item out of range");
        kzg::g_rlc_item = v;
        return 0;
    }
    if (!strcmp(name, "g1fft_split")) {       // takes effect for lanes created afterwards? no: read per call from this global when non-zero
        if (v < 0 || v > KZG_G1FFT_MAX_SPLIT) return set_err(KZGB200_ERR_ARGS, "g1fft_split out of range");
        kzg::g_g1fft_split_override = v;
        return 0;
    }
    if (!strcmp(name, "pairing_lanes")) {
        if (v != 0 && v != 8 && v != 32) return set_err(KZGB200_ERR_ARGS, "pairing_lanes must be 0, 8 or 32");
        kzg::g_pairing_lanes = v;
        return 0;
    }
    if (!strcmp(name, "fiat_shamir")) {
        if (v != 1 && v != 2) return set_err(KZGB200_ERR_ARGS, "fiat_shamir must be 1 or 2");
        kzg::g_fs_variant = v;
        return 0;
    }
    if (!strcmp(name, "vmsm_policy")) {
        if (v < 0 || v > 3) return set_err(KZGB200_ERR_ARGS, "vmsm_policy must be 0..3");
        kzg::g_vmsm_policy = v;
        return 0;
    }
    return set_err(KZGB200_ERR_ARGS, "unknown tunable");
}

// A call that fails half-way (a CUDA error, an allocation that did not fit) returns while kernels and copies it queued may still be
// reading the caller's buffers or the lane's scratch; the public layer calls this before it hands the lane to the next caller.
void lane_quiesce(kzg_lane *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->aux_stream) cudaStreamSynchronize(c->aux_stream);
    for (cudaStream_t q : c->fft_streams) if (q) cudaStreamSynchronize(q);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaGetLastError();
    c->marks_reset();
}

void lane_ctx_free(kzg_lane *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->owns_tables) {
        cudaFree(c->g1_monomial); cudaFree(c->g1_lagrange_brp); cudaFree(c->commit_tab.entries);
        cudaFree(c->fk20_tab.entries); cudaFree(c->roots); cudaFree(c->glv_digits);
        cudaFree(c->pow7); cudaFree(c->ipow7); cudaFree(c->pairing); cudaFree(c->mono64_tab.entries);
    }
    c->msm_partial.release();
    c->in_bytes.release(); c->scalars.release(); c->status.release(); c->sums.release(); c->out_bytes.release();
    c->coeffs.release(); c->cells.release(); c->proofs_xyzz.release(); c->fft_work.release();
    c->in_small.release(); c->in_small2.release(); c->zbuf.release(); c->ybuf.release();
    c->rec_a.release(); c->rec_b.release(); c->rec_meta.release(); c->rec_zev.release(); c->rec_czinv.release();
    c->v_aff1.release(); c->v_aff2.release(); c->v_fr.release(); c->v_meta.release(); c->v_S.release();
    c->v_W.release(); c->v_partial.release(); c->v_in2.release(); c->v_in3.release(); c->v_st2.release();
    c->vm_digits.release(); c->vm_digits256.release(); c->vm_colsum.release(); c->vm_rowdig.release(); c->vm_commsum.release(); c->vm_scratch.release(); c->vm_ws.release(); c->vm_wsb.release(); c->v_pa.release(); c->v_pb.release(); c->v_cst.release(); c->v_st3.release(); c->v_pst.release(); c->v_xst.release(); c->vm_scratch_r.release(); c->vm_ws_r.release(); c->vm_wsb_r.release(); c->ev_cex.release(); c->ev_total.release(); c->ev_index.release();
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_aux_fork) cudaEventDestroy(c->ev_aux_fork);
    if (c->ev_aux_join) cudaEventDestroy(c->ev_aux_join);
    for (cudaStream_t q : c->fft_streams) if (q) cudaStreamDestroy(q);
    for (cudaEvent_t e : c->ev_join) if (e) cudaEventDestroy(e);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : c->ev_piece) if (e) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// streams, events and per-device kernel attributes of a lane on `device`
static int lane_streams_init(kzg_lane *c, int device) {
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    c->device = device;
    if (c->device < 0 || c->device >= ndev) return set_err(KZGB200_ERR_CUDA, "no such CUDA device");
    CU(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    c->sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    {   // the aux stream carries short latency-bound chains that run BESIDE a long pipe-bound kernel of the main stream (second decode of the
        // verifiers, the cell verifier's interpolation chain): highest priority, so that its blocks are placed as soon as slots free up
        int lo_pri = 0, hi_pri = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CU(cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, hi_pri));
    }
    CU(cudaEventCreateWithFlags(&c->ev_aux_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_aux_join, cudaEventDisableTiming));
    CU(cudaEventCreate(&c->ev0));
    CU(cudaEventCreate(&c->ev1));
    for (cudaEvent_t &e : c->ev_piece) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (cudaStream_t &q : c->fft_streams) CU(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    for (cudaEvent_t &e : c->ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    if (const char *e = getenv("KZGB200_G1FFT_SPLIT")) c->g1fft_split = (size_t)std::min(std::max(atoi(e), 1), KZG_G1FFT_MAX_SPLIT);
    if (const char *e = getenv("KZGB200_G1_DENSE_MAX")) c->g1_dense_max = (size_t)std::max(atoi(e), 0);
    if (const char *e = getenv("KZGB200_G1_TWO_LEVEL_MAX")) c->g1_two_level_max = (size_t)std::max(atoi(e), 0);
    if (const char *e = getenv("KZGB200_G1_CHAIN4_MAX")) c->g1_chain4_max = (size_t)std::max(atoi(e), 0);
    return 0;
}

static int ctx_init(kzg_lane *c, const uint8_t *g1m, const uint8_t *g1l, const uint8_t *g2, size_t n_g2, const kzgb200_opts *opts, int device) {
    int rc0 = lane_streams_init(c, device);
    if (rc0) return rc0;
    InitTemps temps;
    c->g2_bytes.assign(g2, g2 + n_g2 * 96);

    int cw = opts && opts->commit_window ? opts->commit_window : 0;
    if (!cw) { const char *e = getenv("KZGB200_COMMIT_WINDOW"); cw = e ? atoi(e) : 13; }
    if (cw < 7 || cw > 15) return set_err(KZGB200_ERR_ARGS, "commit_window must be in 7..15");

    uint8_t *d_in = nullptr; int32_t *d_bad = nullptr;
    CU(cudaMalloc(&d_in, 2 * N_BLOB * 48)); temps.track(d_in);
    CU(cudaMalloc(&d_bad, sizeof(int32_t))); temps.track(d_bad);
    CU(cudaMalloc(&c->g1_monomial, N_BLOB * sizeof(G1Aff)));
    CU(cudaMalloc(&c->g1_lagrange_brp, N_BLOB * sizeof(G1Aff)));
    CU(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), c->stream));
    CU(cudaMemcpyAsync(d_in, g1m, N_BLOB * 48, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(d_in + N_BLOB * 48, g1l, N_BLOB * 48, cudaMemcpyHostToDevice, c->stream));
    k_setup_decompress<<<N_BLOB / 64, 64, 0, c->stream>>>(d_in, c->g1_monomial, N_BLOB, 0, d_bad);
    k_setup_decompress<<<N_BLOB / 64, 64, 0, c->stream>>>(d_in + N_BLOB * 48, c->g1_lagrange_brp, N_BLOB, 12, d_bad);
    c->launches += 2;
    int32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (bad) return set_err(KZGB200_ERR_SETUP, "trusted setup: G1 point failed to decode");

    int rc = build_table(c, c->g1_lagrange_brp, N_BLOB, cw, c->commit_tab);
    if (rc) return rc;

    // Fr twiddles and GLV digit tables
    CU(cudaMalloc(&c->roots, ROOTS_N * sizeof(Fr)));
    CU(cudaMalloc(&c->glv_digits, sizeof(H_GLV_TW)));
    CU(cudaMemcpyAsync(c->glv_digits, H_GLV_TW, sizeof(H_GLV_TW), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyToSymbolAsync(TW_PROG, H_TW_PROG, sizeof(H_TW_PROG), 0, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyToSymbolAsync(DENSE_PROG, H_DENSE_PROG, sizeof(H_DENSE_PROG), 0, cudaMemcpyHostToDevice, c->stream));
    Fr gen; memcpy(gen.v, H_FR_W8192, sizeof gen.v);
    k_init_roots<<<ROOTS_N / 128, 128, 0, c->stream>>>(c->roots, gen);
    CU(cudaMalloc(&c->pow7, 8192 * sizeof(Fr)));
    CU(cudaMalloc(&c->ipow7, 8192 * sizeof(Fr)));
    {
        Fr seven, inv7; memcpy(seven.v, H_FR_SEVEN, sizeof seven.v); memcpy(inv7.v, H_FR_INV7, sizeof inv7.v);
        k_init_pow7<<<8192 / 128, 128, 0, c->stream>>>(c->pow7, c->ipow7, seven, inv7);
        c->launches += 1;
    }
    CU(cudaFuncSetAttribute(k_half_dit_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));
    CU(cudaFuncSetAttribute(k_half_dif_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));
    CU(cudaFuncSetAttribute(k_cells_from_coeffs, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));
    CU(cudaFuncSetAttribute(k_blob_ifft, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));
    CU(cudaFuncSetAttribute(k_coset_fft_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 32));

    // FK20 table (fk20.go:23-52, toeplitz.go:50-93): 64 G1 FFTs of size 128, then the window table
    int fw = opts && opts->fk20_window ? opts->fk20_window : 0;
    if (!fw) { const char *e = getenv("KZGB200_FK20_WINDOW"); fw = e ? atoi(e) : 12; }
    if (fw < 7 || fw > 15) return set_err(KZGB200_ERR_ARGS, "fk20_window must be in 7..15");
    G1 *fk_xyzz = nullptr; G1Aff *fk_aff = nullptr;
    CU(cudaMalloc(&fk_xyzz, 8192 * sizeof(G1))); temps.track(fk_xyzz);
    CU(cudaMalloc(&fk_aff, 8192 * sizeof(G1Aff))); temps.track(fk_aff);
    k_fk20_table_fft<<<64, 64, 0, c->stream>>>(c->g1_monomial, fk_xyzz, c->glv_digits);
    k_to_affine<<<8192 / 128, 128, 0, c->stream>>>(fk_xyzz, fk_aff, 8192);
    c->launches += 3;
    CU(cudaGetLastError());
    rc = build_table(c, fk_aff, 8192, fw, c->fk20_tab);
    if (rc) return rc;

    // verification constants: G2 line tables (GPU, one thread per point) and the 64-point monomial table
    {
        uint8_t h_g2[3 * 96];
        memcpy(h_g2, g2, 96); memcpy(h_g2 + 96, g2 + 96, 96); memcpy(h_g2 + 192, g2 + 64 * 96, 96);
        uint8_t *d_g2 = nullptr; int32_t *d_bad2 = nullptr;
        CU(cudaMalloc(&d_g2, sizeof h_g2)); temps.track(d_g2);
        CU(cudaMalloc(&d_bad2, 4)); temps.track(d_bad2);
        CU(cudaMalloc(&c->pairing, sizeof(PairingConsts)));
        CU(cudaMemsetAsync(d_bad2, 0, 4, c->stream));
        CU(cudaMemcpyAsync(d_g2, h_g2, sizeof h_g2, cudaMemcpyHostToDevice, c->stream));
        k_g2_prepare<<<1, 4, 0, c->stream>>>(d_g2, c->pairing, d_bad2);
        c->launches += 1;
        int32_t bad2 = 0;
        CU(cudaMemcpyAsync(&bad2, d_bad2, 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (bad2) return set_err(KZGB200_ERR_SETUP, "trusted setup: G2 point failed to decode");
    }
    rc = build_table(c, c->g1_monomial, 64, 15, c->mono64_tab);      // 64 points only: 1.8 GB buys 17 instead of 26 windows per scalar
    if (rc) return rc;
    return 0;
}

int lane_ctx_new(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2_monomial, size_t n_g2,
                 const kzgb200_opts *opts, int device, kzg_lane **out) {
    if (!g1_monomial || !g1_lagrange || !g2_monomial || !out) return set_err(KZGB200_ERR_ARGS, "null argument");
    *out = nullptr;
    if (n_g2 < 65) return set_err(KZGB200_ERR_SETUP, "need at least 65 G2 points (api.go:93,115; kzg_multi/kzg_verify.go:92)");
    auto t0 = std::chrono::steady_clock::now();
    kzg_lane *c = new kzg_lane();
    int rc = ctx_init(c, g1_monomial, g1_lagrange, g2_monomial, n_g2, opts, device);
    if (rc) { lane_ctx_free(c); return rc; }
    c->init_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    *out = c;
    return KZGB200_OK;
}

int lane_clone(kzg_lane *f, kzg_lane **out) {
    if (!f || !out) return set_err(KZGB200_ERR_ARGS, "null argument");
    *out = nullptr;
    kzg_lane *c = new kzg_lane();
    c->owns_tables = false;
    int rc = lane_streams_init(c, f->device);
    if (rc) { lane_ctx_free(c); return rc; }
    c->g1_monomial = f->g1_monomial; c->g1_lagrange_brp = f->g1_lagrange_brp; c->g2_bytes = f->g2_bytes;
    c->commit_tab = f->commit_tab; c->fk20_tab = f->fk20_tab; c->mono64_tab = f->mono64_tab;
    c->roots = f->roots; c->glv_digits = f->glv_digits; c->pow7 = f->pow7; c->ipow7 = f->ipow7; c->pairing = f->pairing;
    c->g1fft_split = f->g1fft_split; c->g1_dense_max = f->g1_dense_max; c->g1_two_level_max = f->g1_two_level_max; c->g1_chain4_max = f->g1_chain4_max;
    *out = c;
    return KZGB200_OK;
}

int lane_get_info(kzg_lane *c, kzgb200_info *o) {
    if (!c || !o) return set_err(KZGB200_ERR_ARGS, "null argument");
    memset(o, 0, sizeof *o);
    o->device = c->device; o->sm_count = c->sm_count;
    o->commit_window = c->commit_tab.c; o->commit_windows_per_scalar = c->commit_tab.W;
    o->commit_table_bytes = c->commit_tab.bytes();
    o->fk20_window = c->fk20_tab.c; o->fk20_windows_per_scalar = c->fk20_tab.W;
    o->fk20_table_bytes = c->fk20_tab.bytes();
    o->init_ms = c->init_ms; o->kernel_launches = c->launches;
    return KZGB200_OK;
}

// -------------------------------------------------------------------------------------------
// BlobToKZGCommitment (prove.go:13-34): DeserializeBlob -> MSM against the bit-reversed Lagrange
// SRS -> compress.  Processed in chunks of CHUNK blobs.
// -------------------------------------------------------------------------------------------
static const size_t COMMIT_CHUNK = 4096;

// The 4096-point MSM of m blobs.  One CTA per blob fills the GPU only from about three CTAs per SM upwards; below that
// (the reference's own calling pattern is ONE blob per call, prove.go:13-34) every blob's points are cut into S ranges,
// one CTA each, and a second small kernel adds the S partial sums: 4096 x W gathered additions spread over S x 128
// threads instead of 128, i.e. W + 7 + log2(S) dependent additions per thread instead of 32 W + 7.
static int launch_commit_msm(kzg_lane *c, cudaStream_t st, const uint32_t *scalars, size_t m, const int32_t *d_status, G1 *sums) {
    const int TPB = 128;
    unsigned S = 1;
    while (S < 32 && (size_t)S * m < (size_t)3 * c->sm_count) S <<= 1;
    if (S == 1) {
        launch_msm_fixed(dim3(1, (unsigned)m), TPB, st, scalars, c->commit_tab, N_BLOB, 1, TPB, d_status, sums);
        return 0;
    }
    int rc = c->msm_partial.ensure(m * S * sizeof(G1));
    if (rc) return rc;
    launch_msm_fixed(dim3(S, (unsigned)m), TPB, st, scalars, c->commit_tab, N_BLOB / S, S, TPB, d_status, (G1 *)c->msm_partial.p);
    k_sum_groups<<<(unsigned)m, 32, 32 * sizeof(G1), st>>>((const G1 *)c->msm_partial.p, S, d_status, sums);
    c->launches += 1;
    return 0;
}
#define KZG_H2D_PIECES 4

int lane_blob_to_kzg_commitment(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out48, int32_t *status) {
    if (!c || (n && (!blobs || !out48 || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    const bool in_dev = n && is_device_ptr(blobs), out_dev = n && is_device_ptr(out48), st_dev = n && is_device_ptr(status);
    const size_t chunk = std::min(n, COMMIT_CHUNK);
    if (n == 0) return KZGB200_OK;
    int rc;
    if (!in_dev && (rc = c->in_bytes.ensure(chunk * KZGB200_BYTES_PER_BLOB))) return rc;
    if ((rc = c->scalars.ensure(chunk * N_BLOB * 32))) return rc;
    if ((rc = c->status.ensure(chunk * sizeof(int32_t)))) return rc;
    if ((rc = c->sums.ensure(chunk * sizeof(G1)))) return rc;
    if (!out_dev && (rc = c->out_bytes.ensure(chunk * 48))) return rc;
    const int TPB = 128;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        const uint8_t *d_blobs = blobs + off * KZGB200_BYTES_PER_BLOB;
        int32_t *d_status = st_dev ? status + off : (int32_t *)c->status.p;
        uint8_t *d_out = out_dev ? out48 + off * 48 : (uint8_t *)c->out_bytes.p;
        CU(cudaMemsetAsync(d_status, 0, m * sizeof(int32_t), c->stream));
        // host input: the H2D runs on the copy stream in two pieces and each piece's kernels start as soon as its bytes
        // have landed: a first piece of exactly one wave of MSM CTAs (one CTA per blob, 3 resident per SM), then the
        // rest, whose copy finishes under the first piece's MSM.  (Measured: four equal pieces and geometrically growing
        // pieces both lose more in the partial last waves of their small launches than they hide.)
        // Device input: one launch over the whole chunk.
        size_t bound[KZG_H2D_PIECES + 1] = {0, m, m, m, m};
        if (!in_dev && m >= (size_t)8 * c->sm_count) bound[1] = (size_t)3 * c->sm_count;
        for (size_t pc = 0; pc < KZG_H2D_PIECES && bound[pc] < m; ++pc) {
            const size_t po = bound[pc], pm = bound[pc + 1] - po;
            const uint8_t *pb = d_blobs + po * KZGB200_BYTES_PER_BLOB;
            if (!in_dev) {
                uint8_t *dst = (uint8_t *)c->in_bytes.p + po * KZGB200_BYTES_PER_BLOB;
                CU(cudaMemcpyAsync(dst, pb, pm * KZGB200_BYTES_PER_BLOB, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaEventRecord(c->ev_piece[pc], c->copy_stream));
                CU(cudaStreamWaitEvent(c->stream, c->ev_piece[pc], 0));
                pb = dst;
            }
            c->mark(KZGB200_KC_FR);
            size_t ns = pm * N_BLOB;
            k_blob_to_scalars<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(pb, (uint32_t *)c->scalars.p + po * N_BLOB * 8, d_status + po, ns, N_BLOB);
            c->mark(KZGB200_KC_MSM);
            if ((rc = launch_commit_msm(c, c->stream, (const uint32_t *)c->scalars.p + po * N_BLOB * 8, pm, d_status + po, (G1 *)c->sums.p + po))) return rc;
            c->mark(KZGB200_KC_FINALIZE);
            k_finalize_g1<<<(unsigned)((pm + 64 * KZG_FIN_BATCH - 1) / (64 * KZG_FIN_BATCH)), 64, 0, c->stream>>>((const G1 *)c->sums.p + po, d_out + po * 48, d_status + po, pm, 1);
            c->launches += 3;
        }
        c->mark(-1);
        CU(cudaGetLastError());
        if (!out_dev) CU(cudaMemcpyAsync(out48 + off * 48, d_out, m * 48, cudaMemcpyDeviceToHost, c->stream));
        if (!st_dev) CU(cudaMemcpyAsync(status + off, d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->marks_collect();
    }
    return KZGB200_OK;
}

// -------------------------------------------------------------------------------------------
// ComputeKZGProof (prove.go:85-111) and ComputeBlobKZGProof (prove.go:46-77)
// -------------------------------------------------------------------------------------------
static int open_common(kzg_lane *c, const uint8_t *blobs, const uint8_t *z32, const uint8_t *commitments, size_t n,
                       uint8_t *out_proof, uint8_t *out_y, int32_t *status) {
    if (!c || (n && (!blobs || !out_proof || !status || (!z32 && !commitments)))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    if (n == 0) return KZGB200_OK;
    const bool out_dev = is_device_ptr(out_proof), st_dev = is_device_ptr(status), y_dev = out_y && is_device_ptr(out_y);
    const size_t chunk = std::min(n, COMMIT_CHUNK);
    int rc;
    if ((rc = c->scalars.ensure(chunk * N_BLOB * 32))) return rc;
    if ((rc = c->status.ensure(chunk * sizeof(int32_t)))) return rc;
    if ((rc = c->sums.ensure(chunk * sizeof(G1)))) return rc;
    if ((rc = c->zbuf.ensure(chunk * 32))) return rc;
    if ((rc = c->ybuf.ensure(chunk * 32))) return rc;
    if (!z32 && (rc = c->v_st2.ensure(chunk * sizeof(int32_t)))) return rc;
    if (!out_dev && (rc = c->out_bytes.ensure(chunk * 48))) return rc;
    const int TPB = 128;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        const void *d_aux = nullptr;
        // host blobs travel in geometrically growing pieces on the copy stream (see lane_blob_to_kzg_commitment): every
        // piece runs the whole chain hash -> evaluation -> quotient -> MSM as soon as it has landed.  The hash + evaluation of a
        // piece (latency-bound: one thread per SHA-256) runs on a side stream, under the previous piece's MSM.
        const bool in_dev = is_device_ptr(blobs);
        if (!in_dev && (rc = c->in_bytes.ensure(m * KZGB200_BYTES_PER_BLOB))) return rc;
        {
            // host blobs: the small input travels on the COPY stream ahead of the blob pieces and the main stream waits for it through an event.
            // Queued on the main stream it was observed to complete only after the blob pieces queued behind it on the copy stream, which
            // held back ev_fork and with it every piece's hash (kzgb200_verify.cu: verify_front has the measurement).
            const void *src = z32 ? (const void *)(z32 + off * 32) : (const void *)(commitments + off * 48);
            const size_t bytes = m * (z32 ? 32 : 48);
            if (in_dev || is_device_ptr(src)) { if ((rc = stage_in(c, src, bytes, c->in_small, &d_aux))) return rc; }
            else {
                if ((rc = c->in_small.ensure(bytes))) return rc;
                CU(cudaMemcpyAsync(c->in_small.p, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaEventRecord(c->ev0, c->copy_stream));
                CU(cudaStreamWaitEvent(c->stream, c->ev0, 0));
                d_aux = c->in_small.p;
            }
        }
        int32_t *d_status = st_dev ? status + off : (int32_t *)c->status.p;
        uint8_t *d_out = out_dev ? out_proof + off * 48 : (uint8_t *)c->out_bytes.p;
        uint8_t *d_y = !out_y ? nullptr : y_dev ? out_y + off * 32 : (uint8_t *)c->ybuf.p;
        c->mark(KZGB200_KC_FR);
        CU(cudaMemsetAsync(d_status, 0, m * sizeof(int32_t), c->stream));
        if (d_y) CU(cudaMemsetAsync(d_y, 0, m * 32, c->stream));
        if ((rc = vm_eval_scratch(c, m))) return rc;
        // host input: three pieces (3, then 6 waves of MSM CTAs, then the rest), each one's H2D, hash and evaluation running under the
        // previous piece's MSM (for device input the same split was measured 1.5 % slower than one launch: not used)
        size_t bound[KZG_H2D_PIECES + 1] = {0, m, m, m, m};
        if (!in_dev && m >= (size_t)8 * c->sm_count) {
            bound[1] = (size_t)3 * c->sm_count;
            // tunable "proof_pieces" (2..4): further pieces of 2x, 3x .. the first one.  The evaluation kernels need a whole SM's registers per
            // CTA, so a piece's evaluation only starts when the previous piece's MSM has drained: smaller later pieces expose less of it.
            for (int q = 2; q < g_proof_pieces && q < KZG_H2D_PIECES; ++q) {
                const size_t nxt = bound[q - 1] + (size_t)q * 3 * c->sm_count;
                if (nxt + (size_t)3 * c->sm_count >= m) break;
                bound[q] = nxt;
            }
        }
        const bool split = bound[1] < m;
        if (split) CU(cudaEventRecord(c->ev_fork, c->stream));         // staged aux input and cleared status are visible to the side streams
        for (size_t pc = 0; pc < KZG_H2D_PIECES && bound[pc] < m; ++pc) {
            const size_t po = bound[pc], pm = bound[pc + 1] - po;
            const uint8_t *pb = blobs + (off + po) * KZGB200_BYTES_PER_BLOB;
            cudaStream_t sp = split ? c->fft_streams[pc] : c->stream;
            if (split) CU(cudaStreamWaitEvent(sp, c->ev_fork, 0));
            if (!in_dev) {
                uint8_t *dst = (uint8_t *)c->in_bytes.p + po * KZGB200_BYTES_PER_BLOB;
                CU(cudaMemcpyAsync(dst, pb, pm * KZGB200_BYTES_PER_BLOB, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaEventRecord(c->ev_piece[pc], c->copy_stream));
                CU(cudaStreamWaitEvent(sp, c->ev_piece[pc], 0));
                pb = dst;
            }
            const unsigned gb = (unsigned)((pm + 63) / 64);
            uint32_t *zl = (uint32_t *)c->zbuf.p + po * 8, *ql = (uint32_t *)c->scalars.p + po * N_BLOB * 8;
            c->mark(KZGB200_KC_FR);
            if (z32) {
                k_scalars_from_be<<<gb, 64, 0, sp>>>((const uint8_t *)d_aux + po * 32, zl, d_status + po, pm);
                c->launches += 1;
            } else {
                // The commitment is only validated (decode + subgroup test, the point is discarded: prove.go:56-60), which is ~2.9 ms of one
                // thread's latency whatever the batch, and the hash needs its BYTES only: the check runs on a side stream into a status array
                // of its own, merged below AFTER the blob's own status (DeserializeBlob comes first in the reference, prove.go:52).
                const size_t side = KZG_H2D_PIECES + (pc & 1);      // at most two pieces exist (bound[] above): side streams / events 4 and 5
                cudaStream_t sv = c->fft_streams[side];
                int32_t *d_cst = (int32_t *)c->v_st2.p + po;
                CU(cudaEventRecord(c->ev_piece[side], sp));
                CU(cudaStreamWaitEvent(sv, c->ev_piece[side], 0));
                CU(cudaMemsetAsync(d_cst, 0, pm * sizeof(int32_t), sv));
                k_g1_check<MulCall, 4><<<gb, 64, 0, sv>>>((const uint8_t *)d_aux + po * 48, nullptr, d_cst, pm, 1, 1);
                CU(cudaEventRecord(c->ev_join[side], sv));
                launch_fiat_shamir(sp, pb, (const uint8_t *)d_aux + po * 48, zl, pm);
                c->launches += 2;
            }
            if ((rc = vm_eval_quotient(c, sp, po, pb, zl, d_status + po, ql, d_y ? d_y + po * 32 : nullptr, nullptr, pm))) return rc;
            if (!z32) {
                CU(cudaStreamWaitEvent(sp, c->ev_join[KZG_H2D_PIECES + (pc & 1)], 0));
                k_status_then<<<gb, 64, 0, sp>>>(d_status + po, (const int32_t *)c->v_st2.p + po, pm);
                c->launches += 1;
            }
            if (split) { CU(cudaEventRecord(c->ev_join[pc], sp)); CU(cudaStreamWaitEvent(c->stream, c->ev_join[pc], 0)); }
            c->mark(KZGB200_KC_MSM);
            if ((rc = launch_commit_msm(c, c->stream, ql, pm, d_status + po, (G1 *)c->sums.p + po))) return rc;
            c->mark(KZGB200_KC_FINALIZE);
            k_finalize_g1<<<(unsigned)((pm + 64 * KZG_FIN_BATCH - 1) / (64 * KZG_FIN_BATCH)), 64, 0, c->stream>>>((const G1 *)c->sums.p + po, d_out + po * 48, d_status + po, pm, 1);
            if (d_y) { k_zero_failed<<<gb, 64, 0, c->stream>>>(d_y + po * 32, d_status + po, pm, 32); c->launches += 1; }
            c->launches += 3;
        }
        c->mark(-1);
        CU(cudaGetLastError());
        if (!out_dev) CU(cudaMemcpyAsync(out_proof + off * 48, d_out, m * 48, cudaMemcpyDeviceToHost, c->stream));
        if (out_y && !y_dev) CU(cudaMemcpyAsync(out_y + off * 32, d_y, m * 32, cudaMemcpyDeviceToHost, c->stream));
        if (!st_dev) CU(cudaMemcpyAsync(status + off, d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->marks_collect();
    }
    return KZGB200_OK;
}

int lane_compute_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *z32, size_t n, uint8_t *out_proof48, uint8_t *out_y32, int32_t *status) {
    if (n && (!z32 || !out_y32)) return set_err(KZGB200_ERR_ARGS, "null argument");
    return open_common(c, blobs, z32, nullptr, n, out_proof48, out_y32, status);
}
int lane_compute_blob_kzg_proof(kzg_lane *c, const uint8_t *blobs, const uint8_t *commitments48, size_t n, uint8_t *out48, int32_t *status) {
    if (n && !commitments48) return set_err(KZGB200_ERR_ARGS, "null argument");
    return open_common(c, blobs, nullptr, commitments48, n, out48, nullptr, status);
}

// -------------------------------------------------------------------------------------------
// ComputeCells / ComputeCellsAndKZGProofs (api_eip7594.go:12-52)
// -------------------------------------------------------------------------------------------
static const size_t CELLS_CHUNK = 1024;
#ifndef KZG_FK20_LANES
#define KZG_FK20_LANES 8
#endif

// coefficients (c->coeffs) -> 128 compressed proofs per blob (fk20.go:76-124); buffers must be sized by the caller
// which form of the G1 transform a chunk of m blobs takes, and the points of working storage (c->fft_work) it needs
static inline size_t g1_two_level_max(const kzg_lane *c) { return g_g1_two_level_max >= 0 ? (size_t)g_g1_two_level_max : c->g1_two_level_max; }
static inline size_t g1_chain4_max(const kzg_lane *c) { return g_g1_chain4_max >= 0 ? (size_t)g_g1_chain4_max : c->g1_chain4_max; }
static inline size_t g1fft_work_points(const kzg_lane *c, size_t m) {
    if (m <= c->g1_dense_max) return m * 128 * 65;                  // dense: all 65 x 128 products
    if (m <= g1_two_level_max(c)) return m * 128 * (2 + 16);        // two-level: two working sets + the 16 x 128 products of level I1
    if (m <= g1_chain4_max(c)) return m * 128 * (2 + 4);            // 4 x 4 x 4 x 2: two working sets + 4 x 128 products
    return m * 128;                                                 // staged: the working set
}
static void launch_fk20_proofs(kzg_lane *c, cudaStream_t st, size_t m, const Fr *coeffs, uint32_t *scalars, G1 *sums, G1 *pxyzz,
                               const int32_t *d_status, uint8_t *d_proofs, bool marks) {
    Fr inv128p; memcpy(inv128p.v, H_FR_INV128_PLAIN, sizeof inv128p.v);
    // lanes per 64-point group: 8 for full batches (8 points x W windows per lane, short reduction tree); small batches
    // spread each group over up to 64 lanes so that a single blob occupies 64 CTAs instead of 8
    const int TPB = 128;
    int L = g_fk20_lanes_override ? g_fk20_lanes_override : KZG_FK20_LANES;
    while (L < 64 && (size_t)L * m < (size_t)2 * c->sm_count) L <<= 1;
    // (Running every 128-blob sub-batch's whole chain rows -> MSM -> G1 FFT on a stream of its own, so that MSM blocks fill the tails of FFT
    // stages, was measured twice -- plain streams: 87.4 vs 87.3 ms, all MSMs still run first; priority streams that stagger the chains:
    // 102 vs 86.4 ms -- and removed: profiles/r02_rejected_fk20_overlap.md.)
    k_fk20_rows<<<dim3(64, (unsigned)m), 64, 0, st>>>(coeffs, scalars, d_status, c->roots, inv128p);
    if (marks) c->mark(KZGB200_KC_MSM);
    launch_msm_fixed(dim3(128 / (TPB / L), (unsigned)m), TPB, st, scalars, c->fk20_tab, 64, 128, L, d_status, sums);
    if (marks) c->mark(KZGB200_KC_G1FFT);
    // sums (bit-reversed) --IFFT--> h, keep 64 (toeplitz.go:124), zero-pad (fk20.go:82-85) --FFT--> proofs (bit-reversed);
    // one launch per radix-2 stage, working set in c->fft_work.  The chunk is cut into independent
    // sub-batches on separate streams so that the draining tail of one stage launch (a block is one
    // ~130-doubling scalar multiplication) is filled by another sub-batch's blocks.
    if (m <= c->g1_dense_max) {
        // few blobs: the dense one-level form (g1fft.cuh), ~1.4 ms deep instead of 14 launches of ~1.1 ms
        G1J *prod = (G1J *)c->fft_work.p;       // sized by the caller: max(m * 128, m * 65 * 128) points
        k_g1dense_mul<<<dim3((unsigned)((m * 128 + 31) / 32), 65), 32, 0, st>>>(sums, prod, d_status, (int)m);
        k_g1dense_sum<<<dim3(128, (unsigned)m), 64, 0, st>>>(prod, pxyzz, d_status);
        c->launches += 2;
    } else if (m <= g1_two_level_max(c)) {
        // a medium batch: the two-level form (g1fft.cuh), four twiddle multiplications deep instead of fourteen
        G1J *wa = (G1J *)c->fft_work.p, *wb = wa + m * 128, *prod = wb + m * 128;      // sized by g1fft_work_points
        const unsigned gx = (unsigned)((m + 31) / 32);
        k_g1lvl_mul<0><<<dim3(gx, 128 * 16), 32, 0, st>>>(sums, nullptr, prod, d_status, (int)m, 16);
        k_g1lvl_sum<false><<<(unsigned)((m * 128 + 63) / 64), 64, 0, st>>>(prod, wa, nullptr, d_status, (int)m, 128, 16);
        k_g1lvl_mul<1><<<dim3(gx, 64 * 8), 32, 0, st>>>(nullptr, wa, prod, d_status, (int)m, 8);
        k_g1lvl_sum<false><<<(unsigned)((m * 64 + 63) / 64), 64, 0, st>>>(prod, wb, nullptr, d_status, (int)m, 64, 8);
        k_g1lvl_mul<2><<<dim3(gx, 128 * 8), 32, 0, st>>>(nullptr, wb, prod, d_status, (int)m, 8);
        k_g1lvl_sum<false><<<(unsigned)((m * 128 + 63) / 64), 64, 0, st>>>(prod, wa, nullptr, d_status, (int)m, 128, 8);
        k_g1lvl_mul<3><<<dim3(gx, 128 * 8), 32, 0, st>>>(nullptr, wa, prod, d_status, (int)m, 8);
        k_g1lvl_sum<true><<<(unsigned)((m * 128 + 63) / 64), 64, 0, st>>>(prod, nullptr, pxyzz, d_status, (int)m, 128, 8);
        c->launches += 8;
    } else if (m <= g1_chain4_max(c)) {
        // a medium batch: the 4 x 4 x 4 x 2 form (g1fft.cuh), eight twiddle multiplications deep instead of fourteen
        G1J *wa = (G1J *)c->fft_work.p, *wb = wa + m * 128, *prod = wb + m * 128;      // sized by g1fft_work_points
        const unsigned gx = (unsigned)((m + 31) / 32);
        static const int NOUT[8] = {128, 128, 128, 64, 128, 128, 128, 128}, RR[8] = {4, 4, 4, 2, 2, 4, 4, 2};
#define KZG_LVL8(L, IN, OUT, LASTF) do { \
            k_g1lvl8_mul<L><<<dim3(gx, NOUT[L] * RR[L]), 32, 0, st>>>(sums, IN, prod, d_status, (int)m, RR[L]); \
            k_g1lvl_sum<LASTF><<<(unsigned)((m * NOUT[L] + 63) / 64), 64, 0, st>>>(prod, OUT, pxyzz, d_status, (int)m, NOUT[L], RR[L]); } while (0)
        KZG_LVL8(0, nullptr, wa, false); KZG_LVL8(1, wa, wb, false); KZG_LVL8(2, wb, wa, false); KZG_LVL8(3, wa, wb, false);
        KZG_LVL8(4, wb, wa, false); KZG_LVL8(5, wa, wb, false); KZG_LVL8(6, wb, wa, false); KZG_LVL8(7, wa, nullptr, true);
#undef KZG_LVL8
        c->launches += 16;
    } else {
        const size_t nsplit = std::max<size_t>(1, std::min<size_t>(g_g1fft_split_override ? (size_t)g_g1fft_split_override : c->g1fft_split, (m + 127) / 128));
        const size_t per = (m + nsplit - 1) / nsplit;
        if (nsplit > 1) cudaEventRecord(c->ev_fork, st);
        for (size_t k = 0, off = 0; off < m; ++k, off += per) {
            const int nb = (int)std::min(per, m - off);
            cudaStream_t s = k == 0 ? st : c->fft_streams[k - 1];
            if (k) cudaStreamWaitEvent(s, c->ev_fork, 0);
            const G1 *src = sums + off * 128;
            G1J *work = (G1J *)c->fft_work.p + off * 128;
            G1 *dst = pxyzz + off * 128;
            const int32_t *stt = d_status + off;
            const dim3 grid((unsigned)((nb + KZG_G1FFT_TPB - 1) / KZG_G1FFT_TPB), 64);
#define KZG_STAGE(A, B, C, D, E, F, ...) do { if (g_g1fft_dual) k_g1fft_stage<A, B, C, D, E, F, true><<<grid, KZG_G1FFT_TPB, 0, s>>>(__VA_ARGS__); \
                                                else k_g1fft_stage<A, B, C, D, E, F, false><<<grid, KZG_G1FFT_TPB, 0, s>>>(__VA_ARGS__); } while (0)
            KZG_STAGE(true, true, true, false, false, false, src, work, nullptr, stt, nb, 0);
            for (int lh = 1; lh < 6; ++lh) KZG_STAGE(true, true, false, false, false, false, nullptr, work, nullptr, stt, nb, lh);
            KZG_STAGE(true, true, false, false, true, false, nullptr, work, nullptr, stt, nb, 6);
            KZG_STAGE(false, false, false, false, false, true, nullptr, work, nullptr, stt, nb, 6);
            for (int lh = 5; lh >= 1; --lh) KZG_STAGE(false, false, false, false, false, false, nullptr, work, nullptr, stt, nb, lh);
            KZG_STAGE(false, false, false, true, false, false, nullptr, work, dst, stt, nb, 0);
#undef KZG_STAGE
            c->launches += 14;
            if (k) { cudaEventRecord(c->ev_join[k - 1], s); cudaStreamWaitEvent(st, c->ev_join[k - 1], 0); }
        }
    }
    size_t np = m * 128;
    if (marks) c->mark(KZGB200_KC_FINALIZE);
    k_finalize_g1<<<(unsigned)((np + 64 * KZG_FIN_BATCH - 1) / (64 * KZG_FIN_BATCH)), 64, 0, st>>>(pxyzz, d_proofs, d_status, np, 128);
    c->launches += 3;
}

static int cells_and_proofs(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) {
    if (!c || (n && (!blobs || !out_cells || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    if (n == 0) return KZGB200_OK;
    const bool in_dev = is_device_ptr(blobs), cells_dev = is_device_ptr(out_cells), st_dev = is_device_ptr(status);
    const bool proofs_dev = out_proofs && is_device_ptr(out_proofs);
    const size_t chunk = std::min(n, CELLS_CHUNK);
    int rc;
    if (!in_dev && (rc = c->in_bytes.ensure(chunk * KZGB200_BYTES_PER_BLOB))) return rc;
    if ((rc = c->coeffs.ensure(chunk * N_BLOB * sizeof(Fr)))) return rc;
    if ((rc = c->status.ensure(chunk * sizeof(int32_t)))) return rc;
    if (!cells_dev && (rc = c->cells.ensure(chunk * 262144))) return rc;
    if (out_proofs) {
        if ((rc = c->scalars.ensure(chunk * 8192 * 32))) return rc;
        if ((rc = c->sums.ensure(chunk * 128 * sizeof(G1)))) return rc;
        if ((rc = c->proofs_xyzz.ensure(chunk * 128 * sizeof(G1)))) return rc;
        if ((rc = c->fft_work.ensure(g1fft_work_points(c, chunk) * sizeof(G1J)))) return rc;
        if (!proofs_dev && (rc = c->out_bytes.ensure(chunk * 128 * 48))) return rc;
    }
    Fr inv4096; memcpy(inv4096.v, H_FR_INV4096, sizeof inv4096.v);
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        const uint8_t *d_blobs = blobs + off * KZGB200_BYTES_PER_BLOB;
        int32_t *d_status = st_dev ? status + off : (int32_t *)c->status.p;
        uint8_t *d_cells = cells_dev ? out_cells + off * 262144 : (uint8_t *)c->cells.p;
        uint8_t *d_proofs = !out_proofs ? nullptr : proofs_dev ? out_proofs + off * 6144 : (uint8_t *)c->out_bytes.p;
        // host input: H2D in pieces on the copy stream, the per-blob Fr kernels of a piece start when
        // its bytes have landed.  The cells are final after k_coset_fft_cells: their D2H (the bulk of
        // the output bytes) also runs on the copy stream, underneath the FK20 kernels.
        const size_t pieces = in_dev ? 1 : std::min<size_t>(KZG_H2D_PIECES, (m + 63) / 64);
        const size_t per = (m + pieces - 1) / pieces;
        c->mark(KZGB200_KC_FR);
        for (size_t pc = 0, po = 0; po < m; ++pc, po += per) {
            size_t pm = std::min(per, m - po);
            const uint8_t *pb = d_blobs + po * KZGB200_BYTES_PER_BLOB;
            if (!in_dev) {
                uint8_t *dst = (uint8_t *)c->in_bytes.p + po * KZGB200_BYTES_PER_BLOB;
                CU(cudaMemcpyAsync(dst, pb, pm * KZGB200_BYTES_PER_BLOB, cudaMemcpyHostToDevice, c->copy_stream));
                CU(cudaEventRecord(c->ev_piece[pc], c->copy_stream));
                CU(cudaStreamWaitEvent(c->stream, c->ev_piece[pc], 0));
                pb = dst;
            }
            k_blob_ifft<<<(unsigned)pm, KZG_NTT_THREADS, 4096 * 32, c->stream>>>(pb, (Fr *)c->coeffs.p + po * N_BLOB, d_status + po, c->roots, inv4096);
            k_coset_fft_cells<<<(unsigned)pm, KZG_NTT_THREADS, 4096 * 32, c->stream>>>((const Fr *)c->coeffs.p + po * N_BLOB, pb, d_cells + po * 262144, d_status + po, c->roots);
            c->launches += 2;
        }
        if (!cells_dev) {
            CU(cudaEventRecord(c->ev0, c->stream));
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev0, 0));
            CU(cudaMemcpyAsync(out_cells + off * 262144, d_cells, m * 262144, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->ev1, c->copy_stream));
        }
        if (out_proofs) launch_fk20_proofs(c, c->stream, m, (const Fr *)c->coeffs.p, (uint32_t *)c->scalars.p, (G1 *)c->sums.p, (G1 *)c->proofs_xyzz.p, d_status, d_proofs, true);
        c->mark(-1);
        CU(cudaGetLastError());
        if (!cells_dev) CU(cudaStreamWaitEvent(c->stream, c->ev1, 0));      // the staging buffer is reused by the next chunk
        if (out_proofs && !proofs_dev) CU(cudaMemcpyAsync(out_proofs + off * 6144, d_proofs, m * 6144, cudaMemcpyDeviceToHost, c->stream));
        if (!st_dev) CU(cudaMemcpyAsync(status + off, d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->marks_collect();
    }
    return KZGB200_OK;
}

int lane_compute_cells(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out_cells, int32_t *status) {
    return cells_and_proofs(c, blobs, n, out_cells, nullptr, status);
}
int lane_compute_cells_and_kzg_proofs(kzg_lane *c, const uint8_t *blobs, size_t n, uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) {
    if (n && !out_proofs) return set_err(KZGB200_ERR_ARGS, "null argument");
    return cells_and_proofs(c, blobs, n, out_cells, out_proofs, status);
}

// -------------------------------------------------------------------------------------------
// RecoverCellsAndComputeKZGProofs / RecoverCells (api_eip7594.go:93-161, api_eip.go:8-15)
// cell_ids and counts are read on the host (validation order of api_eip7594.go:93-113).
// -------------------------------------------------------------------------------------------
static const size_t RECOVER_CHUNK = 1024;

int lane_recover_cells_and_kzg_proofs(kzg_lane *c, const uint64_t *cell_ids, const uint64_t *counts, const uint8_t *cells, size_t n,
                                         uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) {
    if (!c || (n && (!counts || !out_cells || !status))) return set_err(KZGB200_ERR_ARGS, "null argument");
    if (n && (is_device_ptr(cell_ids) || is_device_ptr(counts))) return set_err(KZGB200_ERR_ARGS, "cell_ids/counts must be host pointers");
    CU(cudaSetDevice(c->device));
    c->timing_reset();
    if (n == 0) return KZGB200_OK;
    const bool cells_in_dev = cells && is_device_ptr(cells), cells_dev = is_device_ptr(out_cells), st_dev = is_device_ptr(status);
    const bool proofs_dev = out_proofs && is_device_ptr(out_proofs);
    // host-side validation + metadata
    std::vector<int32_t> h_status(n, KZGB200_OK), h_slot(n * 128, -1);
    std::vector<uint8_t> h_present(n * 128, 0);
    std::vector<uint64_t> h_base(n, 0);
    uint64_t base = 0;
    for (size_t b = 0; b < n; ++b) {
        const uint64_t *ids = cell_ids + base;
        uint64_t cnt = counts[b];
        h_base[b] = base;
        int st = KZGB200_OK;
        for (uint64_t i = 1; i < cnt && !st; ++i) if (ids[i] <= ids[i - 1]) st = KZGB200_CELL_IDS_NOT_ASCENDING;
        for (uint64_t i = 0; i < cnt && !st; ++i) if (ids[i] >= 128) st = KZGB200_BAD_CELL_INDEX;
        if (!st && cnt < 64) st = KZGB200_NOT_ENOUGH_CELLS;
        if (!st) for (uint64_t i = 0; i < cnt; ++i) { h_present[b * 128 + ids[i]] = 1; h_slot[b * 128 + ids[i]] = (int32_t)i; }
        h_status[b] = st;
        base += cnt;
    }
    const uint64_t total_cells = base;
    if (total_cells && !cells) return set_err(KZGB200_ERR_ARGS, "null argument");
    const size_t chunk = std::min(n, RECOVER_CHUNK);
    int rc;
    if ((rc = c->rec_a.ensure(chunk * 8192 * sizeof(Fr)))) return rc;
    if ((rc = c->rec_b.ensure(chunk * 8192 * sizeof(Fr)))) return rc;
    if ((rc = c->rec_zev.ensure(chunk * 128 * sizeof(Fr)))) return rc;
    if ((rc = c->rec_czinv.ensure(chunk * 128 * sizeof(Fr)))) return rc;
    if ((rc = c->rec_meta.ensure(chunk * (128 * 4 + 128 + 8)))) return rc;
    if ((rc = c->coeffs.ensure(chunk * N_BLOB * sizeof(Fr)))) return rc;
    if ((rc = c->status.ensure(chunk * sizeof(int32_t)))) return rc;
    if (!cells_dev && (rc = c->cells.ensure(chunk * 262144))) return rc;
    if (out_proofs) {
        if ((rc = c->scalars.ensure(chunk * 8192 * 32))) return rc;
        if ((rc = c->sums.ensure(chunk * 128 * sizeof(G1)))) return rc;
        if ((rc = c->proofs_xyzz.ensure(chunk * 128 * sizeof(G1)))) return rc;
        if ((rc = c->fft_work.ensure(g1fft_work_points(c, chunk) * sizeof(G1J)))) return rc;
        if (!proofs_dev && (rc = c->out_bytes.ensure(chunk * 128 * 48))) return rc;
    }
    Fr inv8192; memcpy(inv8192.v, H_FR_INV8192, sizeof inv8192.v);
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        // cells of this chunk are contiguous in the flat input
        uint64_t c0 = h_base[off], c1 = off + m < n ? h_base[off + m] : total_cells;
        const uint8_t *d_cells_in = nullptr;
        if (c1 > c0) {
            if (cells_in_dev) d_cells_in = cells + c0 * 2048;
            else {
                if ((rc = c->in_bytes.ensure((c1 - c0) * 2048))) return rc;
                CU(cudaMemcpyAsync(c->in_bytes.p, cells + c0 * 2048, (c1 - c0) * 2048, cudaMemcpyHostToDevice, c->stream));
                d_cells_in = (const uint8_t *)c->in_bytes.p;
            }
        }
        int32_t *d_slot = (int32_t *)c->rec_meta.p;
        uint8_t *d_present = (uint8_t *)(d_slot + chunk * 128);
        uint64_t *d_base = (uint64_t *)(d_present + chunk * 128);
        std::vector<uint64_t> rel(m);
        for (size_t b = 0; b < m; ++b) rel[b] = h_base[off + b] - c0;
        int32_t *d_status = st_dev ? status + off : (int32_t *)c->status.p;
        CU(cudaMemcpyAsync(d_slot, h_slot.data() + off * 128, m * 128 * 4, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_present, h_present.data() + off * 128, m * 128, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_base, rel.data(), m * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(d_status, h_status.data() + off, m * 4, cudaMemcpyHostToDevice, c->stream));
        uint8_t *d_cells = cells_dev ? out_cells + off * 262144 : (uint8_t *)c->cells.p;
        uint8_t *d_proofs = !out_proofs ? nullptr : proofs_dev ? out_proofs + off * 6144 : (uint8_t *)c->out_bytes.p;
        Fr *A = (Fr *)c->rec_a.p, *B = (Fr *)c->rec_b.p;
        c->mark(KZGB200_KC_FR);
        k_rec_vanishing<<<(unsigned)m, 128, 0, c->stream>>>(d_present, d_status, c->roots, c->pow7, (Fr *)c->rec_zev.p, (Fr *)c->rec_czinv.p);
        k_rec_scale<<<dim3(8192 / 256, (unsigned)m), 256, 0, c->stream>>>(d_cells_in, d_slot, d_base, (const Fr *)c->rec_zev.p, d_status, A);
        k_half_dit_inv<<<dim3(2, (unsigned)m), KZG_NTT_THREADS, 4096 * 32, c->stream>>>(A, B, d_status, c->roots);
        k_inv_combine<<<dim3(4096 / 256, (unsigned)m), 256, 0, c->stream>>>(B, A, d_status, c->roots, c->pow7, inv8192, 1);
        k_half_dif_fwd<<<dim3(2, (unsigned)m), KZG_NTT_THREADS, 4096 * 32, c->stream>>>(A, B, d_status, c->roots, (const Fr *)c->rec_czinv.p);
        k_half_dit_inv<<<dim3(2, (unsigned)m), KZG_NTT_THREADS, 4096 * 32, c->stream>>>(B, A, d_status, c->roots);
        k_inv_combine<<<dim3(4096 / 256, (unsigned)m), 256, 0, c->stream>>>(A, (Fr *)c->coeffs.p, d_status, c->roots, c->ipow7, inv8192, 0);
        k_cells_from_coeffs<<<dim3(2, (unsigned)m), KZG_NTT_THREADS, 4096 * 32, c->stream>>>((const Fr *)c->coeffs.p, d_cells, d_status, c->roots);
        c->launches += 8;
        // the cells are final here: their D2H (the bulk of the output bytes) runs on the copy stream underneath the FK20 kernels, as in cells_and_proofs
        if (!cells_dev) {
            CU(cudaEventRecord(c->ev0, c->stream));
            CU(cudaStreamWaitEvent(c->copy_stream, c->ev0, 0));
            CU(cudaMemcpyAsync(out_cells + off * 262144, d_cells, m * 262144, cudaMemcpyDeviceToHost, c->copy_stream));
            CU(cudaEventRecord(c->ev1, c->copy_stream));
        }
        if (out_proofs) launch_fk20_proofs(c, c->stream, m, (const Fr *)c->coeffs.p, (uint32_t *)c->scalars.p, (G1 *)c->sums.p, (G1 *)c->proofs_xyzz.p, d_status, d_proofs, true);
        c->mark(-1);
        CU(cudaGetLastError());
        if (!cells_dev) CU(cudaStreamWaitEvent(c->stream, c->ev1, 0));      // (also: the staging buffer is reused by the next chunk)
        if (out_proofs && !proofs_dev) CU(cudaMemcpyAsync(out_proofs + off * 6144, d_proofs, m * 6144, cudaMemcpyDeviceToHost, c->stream));
        if (!st_dev) CU(cudaMemcpyAsync(status + off, d_status, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->marks_collect();
    }
    return KZGB200_OK;
}

// -------------------------------------------------------------------------------------------
// debug entry points (include/kzgb200_debug.h)
// -------------------------------------------------------------------------------------------
static int dbg_binop(int which, const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op, int limbs) {
    uint32_t *da, *db, *dout;
    size_t bytes = (size_t)n * limbs * 4;
    CU(cudaMalloc(&da, bytes)); CU(cudaMalloc(&db, bytes)); CU(cudaMalloc(&dout, bytes));
    CU(cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice));
    if (which == 0) k_dbg_fp_mul<<<(n + 63) / 64, 64>>>(da, db, dout, n, op);
    else k_dbg_fr_mul<<<(n + 63) / 64, 64>>>(da, db, dout, n, op);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}
int kzgb200_dbg_fp_op(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op) { return dbg_binop(0, a, b, out, n, op, 12); }
int kzgb200_dbg_fr_op(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op) { return dbg_binop(1, a, b, out, n, op, 8); }
int kzgb200_dbg_g1_op(const uint8_t *a48, const uint8_t *b48, uint8_t *out48, int n, int op) {
    uint8_t *da, *db, *dout;
    size_t bytes = (size_t)n * 48;
    CU(cudaMalloc(&da, bytes)); CU(cudaMalloc(&db, bytes)); CU(cudaMalloc(&dout, bytes));
    CU(cudaMemcpy(da, a48, bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(db, b48, bytes, cudaMemcpyHostToDevice));
    k_dbg_g1_op<<<(n + 31) / 32, 32>>>(da, db, dout, n, op);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out48, dout, bytes, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return 0;
}
// measured device-wide 32-bit IMAD rate (results/s); the roofline denominator of SURVEY 8(d)
int kzgb200_bench_imad(int device, int mode, double *imad_per_s, double *ms_out) {
    CU(cudaSetDevice(device));
    cudaDeviceProp prop; CU(cudaGetDeviceProperties(&prop, device));
    int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    if (mode >= 3) { threads = 128; iters = 512; }
    uint32_t *d; CU(cudaMalloc(&d, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    double best = 1e30;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(e0));
        if (mode == 0) k_imad_peak<0><<<blocks, threads>>>(d, iters, 17u + rep);
        else if (mode == 1) k_imad_peak<1><<<blocks, threads>>>(d, iters, 17u + rep);
        else if (mode == 2) k_imad_peak<2><<<blocks, threads>>>(d, iters, 17u + rep);
        else if (mode == 3) k_fp_chain<false><<<blocks, threads>>>(d, iters, 17u + rep);
        else k_fp_chain<true><<<blocks, threads>>>(d, iters, 17u + rep);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms; CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    double ops = (double)blocks * threads * iters * (mode >= 3 ? 1.0 : 8.0 * 8.0);   // 8 unrolled x 8 mads; modes 3 / 4: field operations
    *imad_per_s = ops / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaFree(d); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

}  // extern "C"
