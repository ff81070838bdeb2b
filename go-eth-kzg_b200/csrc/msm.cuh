// Fixed-base multi-scalar multiplication against full digit tables resident in HBM.
//
// Replaces internal/multiexp/multiexp.go:20-26 (-> gnark MultiExp, generic Pippenger) for the two
// fixed bases of the hot path: the bit-reversed Lagrange SRS (internal/kzg/srs.go:56-62,
// api.go:131) and the FK20 table (internal/kzg_multi/fk20/toeplitz.go:87,113-119).
//
// B200-first design: the bases never change, HBM is 180 GB, and a batch holds thousands of
// blobs, so instead of Pippenger buckets (whose per-blob bucket state, 4096 x 192 B, fits no
// shared memory and costs a 2*2^(c-1)-add reduction per blob) the table stores EVERY signed
// digit multiple of every window of every base point in affine form:
//     table[j*row + rowoff[k] + (d-1)] = d * 2^(bitpos[k]) * P_j      d = 1..2^(bits[k]-1)   (see MsmTable::plan)
// and an MSM is just n*W mixed additions of gathered table entries plus a log-depth reduction --
// no buckets, no sorting, no atomics, no doublings.  c = 13 gives 32 GB for the 4096-point base.
// Entries are 96-byte affine points (3 x 32 B sectors per gather); the gathers are random over
// the table, so the kernel is IMAD-bound with ~100-300 GB/s of scattered HBM reads in flight.
#pragma once
#include "codec.cuh"
#include <cstdlib>

namespace kzg {

// Window plan: W windows covering exactly 256 bits.  With c the requested (minimum) window size,
// W = floor(256/c) and the top 256 - W*c windows are one bit wider, which saves the nearly empty
// extra window of a uniform split (c = 15: 16 x 15 + 1 x 16 bits = 17 windows instead of 18 at the
// same table size; c = 14: 14 x 14 + 4 x 15 = 18 instead of 19).  Falls back to ceil(256/c)
// uniform windows when the remainder does not fit.
#define KZG_MAX_WINDOWS 40
struct MsmTable {
    G1Aff *entries;     // device
    int npts, c, W;
    uint32_t row_entries;                 // entries per base point = sum_k 2^(bits[k]-1)
    uint8_t bits[KZG_MAX_WINDOWS];        // window sizes
    uint16_t bitpos[KZG_MAX_WINDOWS];     // first bit of window k
    uint32_t rowoff[KZG_MAX_WINDOWS];     // entry offset of window k inside a point's row
    size_t bytes() const { return (size_t)npts * row_entries * sizeof(G1Aff); }
    void plan(int npts_, int c_) {
        npts = npts_; c = c_;
        int Wf = 256 / c, rem = 256 - Wf * c;
        if (rem > Wf) { W = Wf + 1; rem = 0; } else W = Wf;
        uint32_t off = 0; int pos = 0;
        for (int k = 0; k < W; ++k) {
            int b = c + ((rem && k >= W - rem) ? 1 : 0);
            bits[k] = (uint8_t)b; bitpos[k] = (uint16_t)pos; rowoff[k] = off;
            off += 1u << (b - 1); pos += b;
        }
        row_entries = off;
    }
};

// signed-digit recoding state for one scalar (plain little-endian limbs, < 2^255)
struct DigitStream {
    uint32_t s[8];
    uint32_t carry;
    __device__ __forceinline__ void init(const uint32_t *limbs) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s[i] = limbs[i];
        carry = 0;
    }
    // digit of the window [bit, bit+c) (must be called window by window from the bottom); signed digit
    // in [-(2^(c-1) - 1), 2^(c-1)].  The top window never overflows because bit 255 of a scalar is 0.
    __device__ __forceinline__ int next(int bit, int c) {
        int wi = bit >> 5, sh = bit & 31;
        // dynamic limb index without local memory: select via unrolled compare
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { if (i == wi) lo = s[i]; if (i == wi + 1) hi = s[i]; }
        uint32_t raw = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh) & ((1u << c) - 1u);
        raw += carry;
        int H = 1 << (c - 1);
        if ((int)raw > H) { carry = 1; return (int)raw - (1 << c); }
        carry = 0;
        return (int)raw;
    }
};

__device__ __forceinline__ G1Aff load_aff(const G1Aff *p) {
    G1Aff r;
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 t[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) t[i] = __ldg(q + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        r.x.v[4 * i] = t[i].x; r.x.v[4 * i + 1] = t[i].y; r.x.v[4 * i + 2] = t[i].z; r.x.v[4 * i + 3] = t[i].w;
        r.y.v[4 * i] = t[3 + i].x; r.y.v[4 * i + 1] = t[3 + i].y; r.y.v[4 * i + 2] = t[3 + i].z; r.y.v[4 * i + 3] = t[3 + i].w;
    }
    return r;
}

// Generic launch shape: blockIdx.y = blob; each block of blockDim.x threads serves
// blockDim.x / L groups of `group_pts` consecutive base points, L lanes cooperating per group.
// Scalars: [blob][n_groups*group_pts][8] plain limbs (table order).  out: [blob][n_groups] XYZZ.
//   commit:  group_pts = 4096, n_groups = 1,   L = blockDim.x
//   FK20:    group_pts = 64,   n_groups = 128, L = 8 (16 groups per 128-thread block)
//
// Variant V (bit mask; chosen at run time by msm_variant(), every variant is bit-exact):
//   bit 0  field products through ONE out-of-line copy (MulCall) instead of ten inlined ones: the inlined mixed addition
//          is ~70 KB of SASS, and ncu shows what that costs -- sm__icc_request_hit_rate 73 %, `no_instruction` the second
//          largest stall (2.8 warps per issue) -- while the call costs ~36 register moves on the idle ALU pipe
//   bit 1  the gather is STAGED: the 96-byte entry of item i+1 travels global -> shared with cp.async (no registers)
//          while the ten products of item i run, so the random-HBM latency (~1.5 us) is off the critical path; the two
//          96-byte stages per thread alias the 192 bytes per thread the reduction tree needs afterwards
//   bit 2  Y3 = R (Q - X3) - Y1 PPP under one Montgomery reduction (Mont::mul_add_mul): 2 604 instead of 2 748 wide products
//          per gathered addition
//   bit 3  the accumulator lives in shared memory (g1_add_affine_smem) and the kernel is held to 128 registers: four CTAs per
//          SM instead of three
extern __shared__ unsigned char msm_smem[];

template <int V> struct MsmMul { typedef MulInline type; };
template <> struct MsmMul<1> { typedef MulCall type; };
template <> struct MsmMul<4> { typedef MulInlineLazy type; };
template <> struct MsmMul<5> { typedef MulCallLazy type; };

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Fp at shared-memory slot `comp` of thread t: layout [component][16-byte piece][thread] (conflict-free 16-byte accesses)
struct SmemAcc {
    uint4 *base; int nthr, t;
    __device__ __forceinline__ Fp ld(int comp) const {
        uint4 a = base[(comp * 3 + 0) * nthr + t], b = base[(comp * 3 + 1) * nthr + t], c = base[(comp * 3 + 2) * nthr + t];
        Fp r;
        r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
        r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
        return r;
    }
    __device__ __forceinline__ void st(int comp, const Fp &r) const {
        base[(comp * 3 + 0) * nthr + t] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
        base[(comp * 3 + 1) * nthr + t] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
        base[(comp * 3 + 2) * nthr + t] = make_uint4(r.v[8], r.v[9], r.v[10], r.v[11]);
    }
    __device__ __forceinline__ G1 load() const { G1 r; r.X = ld(0); r.Y = ld(1); r.ZZ = ld(2); r.ZZZ = ld(3); return r; }
    __device__ __forceinline__ void store(const G1 &r) const { st(0, r.X); st(1, r.Y); st(2, r.ZZ); st(3, r.ZZZ); }
};
// acc += b with the accumulator resident in shared memory: every coordinate is loaded where it is used and stored as soon
// as it is final, so the 48 registers of a register-resident accumulator are free for a fourth CTA per SM (variant bit 3)
template <class M_> __device__ __forceinline__ void g1_add_affine_smem(const SmemAcc &A, const G1Aff &b) {
    Fp ZZ = A.ld(2);
    if (ZZ.is_zero()) { A.st(0, b.x); A.st(1, b.y); A.st(2, Fp::one()); A.st(3, Fp::one()); return; }
    Fp U2 = M_::mul(b.x, ZZ);
    Fp S2 = M_::mul(b.y, A.ld(3));
    Fp Pd = Fp::sub(U2, A.ld(0));
    Fp R = Fp::sub(S2, A.ld(1));
    if (Pd.is_zero()) {
        G1 acc = A.load();
        if (R.is_zero()) g1_dbl_affine(&acc, &b);
        else acc = G1::infinity();
        A.store(acc);
        return;
    }
    Fp PP = M_::sqr(Pd);
    Fp PPP = M_::mul(Pd, PP);
    Fp Q = M_::mul(A.ld(0), PP);
    Fp X3 = Fp::sub(Fp::sub(M_::sqr(R), PPP), Fp::dbl(Q));
    A.st(0, X3);
    A.st(1, M_::msub(R, Fp::sub(Q, X3), A.ld(1), PPP));
    A.st(2, M_::mul(A.ld(2), PP));
    A.st(3, M_::mul(A.ld(3), PPP));
}

template <int V>
static __global__ void __launch_bounds__(128, (V & 8) ? 4 : 3) k_msm_fixed(const uint32_t *__restrict__ scalars, MsmTable tab, int group_pts,
                                                    int n_groups, int L, const int32_t *__restrict__ status, G1 *__restrict__ out) {
    typedef typename MsmMul<(V & 5)>::type M_;
    const int blob = blockIdx.y, t = threadIdx.x;
    if (status && status[blob] != ST_OK) return;
    const int lane = t & (L - 1);
    const int group = blockIdx.x * (blockDim.x / L) + t / L;
    const int total = group_pts * n_groups;
    G1 acc = G1::infinity();
    // shared memory: [reduction tree / staging: blockDim * 192 B] [variant bit 3: accumulators, blockDim * 192 B]
    SmemAcc A;
    A.base = reinterpret_cast<uint4 *>(msm_smem) + 12 * blockDim.x; A.nthr = blockDim.x; A.t = t;
    if (V & 8) A.store(acc);
    if (group < n_groups) {
        const uint32_t *sc = scalars + ((size_t)blob * total + (size_t)group * group_pts) * 8;
        if (V & 2) {
            // staged gather: items (point j, window k) in order; the generator runs one item ahead of the consumer.
            // Shared layout [stage][16-byte piece][thread]: consecutive threads hit consecutive 16-byte words.
            uint4 *stage = reinterpret_cast<uint4 *>(msm_smem);
            const int nthr = blockDim.x;
            const int n_items = ((group_pts - lane + L - 1) / L) * tab.W;
            DigitStream ds;
            int gj = lane - L, gk = tab.W;          // generator position
            const G1Aff *row = nullptr;
            int dcur = 0, dnext = 0;
            auto issue = [&](int buf) -> int {       // advances the generator, issues the copy of its item, returns the digit
                if (gk == tab.W) {
                    gj += L; gk = 0;
                    const uint4 *q = reinterpret_cast<const uint4 *>(sc + (size_t)gj * 8);
                    uint4 a = __ldg(q), b = __ldg(q + 1);
                    uint32_t l[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    ds.init(l);
                    row = tab.entries + (size_t)(group * group_pts + gj) * tab.row_entries;
                }
                int d = ds.next(tab.bitpos[gk], tab.bits[gk]);
                if (d) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(row + tab.rowoff[gk] + ((d < 0 ? -d : d) - 1));
#pragma unroll
                    for (int i = 0; i < 6; ++i) cp_async16(stage + (buf * 6 + i) * nthr + t, src + i);
                }
                ++gk;
                return d;
            };
            if (n_items > 0) dcur = issue(0);
            cp_async_commit();
            for (int it = 0, buf = 0; it < n_items; ++it, buf ^= 1) {
                if (it + 1 < n_items) dnext = issue(buf ^ 1);
                cp_async_commit();
                cp_async_wait<1>();
                if (dcur) {
                    G1Aff e;
                    uint4 w[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = stage[(buf * 6 + i) * nthr + t];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        e.x.v[4 * i] = w[i].x; e.x.v[4 * i + 1] = w[i].y; e.x.v[4 * i + 2] = w[i].z; e.x.v[4 * i + 3] = w[i].w;
                        e.y.v[4 * i] = w[3 + i].x; e.y.v[4 * i + 1] = w[3 + i].y; e.y.v[4 * i + 2] = w[3 + i].z; e.y.v[4 * i + 3] = w[3 + i].w;
                    }
                    if (dcur < 0) e.y = Fp::neg(e.y);
                    if (V & 8) g1_add_affine_smem<M_>(A, e); else g1_add_affine<M_>(acc, e);
                }
                dcur = dnext;
            }
            cp_async_wait<0>();
        } else {
            for (int j = lane; j < group_pts; j += L) {
                DigitStream ds;
                {
                    const uint4 *q = reinterpret_cast<const uint4 *>(sc + (size_t)j * 8);
                    uint4 a = __ldg(q), b = __ldg(q + 1);
                    uint32_t l[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    ds.init(l);
                }
                const G1Aff *row = tab.entries + (size_t)(group * group_pts + j) * tab.row_entries;
                for (int k = 0; k < tab.W; ++k) {
                    int d = ds.next(tab.bitpos[k], tab.bits[k]);
                    if (d == 0) continue;
                    int mag = d < 0 ? -d : d;
                    G1Aff e = load_aff(row + tab.rowoff[k] + (mag - 1));
                    if (d < 0) e.y = Fp::neg(e.y);
                    if (V & 8) g1_add_affine_smem<M_>(A, e); else g1_add_affine<M_>(acc, e);
                }
            }
        }
    }
    if (V & 8) acc = A.load();
    // tree reduction over the L lanes of each group through shared memory (the staging area is dead by now)
    if (V & 2) __syncthreads();
    G1 *sm = reinterpret_cast<G1 *>(msm_smem);
    sm[t] = acc;
    __syncthreads();
    // compacted: the (groups per block) * s additions of a level run on the FIRST that many threads, so whole warps go idle
    // instead of every warp running the 14-product addition for half, a quarter, an eighth of its lanes
    const int gpb = blockDim.x / L;
    for (int s = L >> 1; s > 0; s >>= 1) {
        if (t < gpb * s) {
            const int sh = __ffs(s) - 1, g = t >> sh, l = t & (s - 1);
            g1_add_ool(&sm[g * L + l], &sm[g * L + l + s]);
        }
        __syncthreads();
    }
    if (lane == 0 && group < n_groups) out[(size_t)blob * n_groups + group] = sm[t];
}

// run-time choice of the variant: KZGB200_MSM_VARIANT (0..7) overrides the default; translation units that define
// KZG_MSM_ALL_VARIANTS (kzgb200.cu: the proving paths) carry all eight, the others only the default
#ifndef KZG_MSM_DEFAULT_VARIANT
#define KZG_MSM_DEFAULT_VARIANT 13   /* measured on B200 (scripts/msm_variants.py, profiles/r02_msm_variants.md) */
#endif
inline int g_msm_variant_override = -1;       // kzgb200_dbg_set_tunable("msm_variant", v) (experiments: one context, every variant)
inline int g_vmsm_policy = 1;                 // kzgb200_dbg_set_tunable("vmsm_policy", 0..3): see k_vmsm_buckets (vmsm.cuh)
inline int g_g1fft_dual = 0;                  // kzgb200_dbg_set_tunable("g1fft_dual", 0 | 1): G1 FFT stage kernel with every pair of independent field products in one dual body, 2 CTAs/SM (g1fft.cuh: jac_mul_prog_dual_at); measured SLOWER (33.8 -> 34.5 ms), off
inline int g_decode_dual = 1;                 // kzgb200_dbg_set_tunable("decode_dual", 0 | 1): the same in the subgroup test of k_g1_check (vm_g1_check); measured 29.7 -> 29.4 ms per 528 k points
inline int g_large_window = 4;                // kzgb200_dbg_set_tunable("large_window", 4 | 8): window bits of the large verdicts' column MSMs (kzgb200_verify.cu)
inline int g_large_item = 0;                  // kzgb200_dbg_set_tunable("large_item", n): run length of their work items; 0 = default (128 / 512)
inline int g_optimistic = 1;                  // kzgb200_dbg_set_tunable("optimistic", 0 | 1): see lane_verify_cell_kzg_proof_batch (KZGB200_OPTIMISTIC=0 also turns it off)
inline int g_verify_overlap = 1;              // kzgb200_dbg_set_tunable("verify_overlap", 0 | 1): cell verifier's interpolation chain on the aux stream beside the proofs' decode
inline int g_decode_minb = 4;                 // kzgb200_dbg_set_tunable("decode_minb", 4 | 6 | 8): see vm_g1_check (kzgb200_vmsm.cu)
inline int g_g1_two_level_max = -1;           // kzgb200_dbg_set_tunable("g1_two_level_max", n): chunks of up to n blobs take the two-level G1 transform (g1fft.cuh); -1 = context default, 0 = never
inline int g_g1_chain4_max = -1;              // kzgb200_dbg_set_tunable("g1_chain4_max", n): chunks of up to n blobs (above g1_two_level_max) take the 4 x 4 x 4 x 2 G1 transform; -1 = context default, 0 = never
inline int g_proof_pieces = 3;                // kzgb200_dbg_set_tunable("proof_pieces", 2..4): H2D pieces of the host-buffer ComputeKZGProof / ComputeBlobKZGProof paths; measured per 4096 blobs: 2 -> 118.4, 3 -> 109.3, 4 -> 110.3 ms
inline int g_tail_split = 0;                  // kzgb200_dbg_set_tunable("tail_split", 0 | 1): see kzgb200_verify_cell_kzg_proof_batch (kzgb200_api.cu).  MEASURED: no gain (47.35 vs 47.2 ms), default off
inline int g_tail_min_cells = 128 << 10;        // kzgb200_dbg_set_tunable("tail_min_cells", n): smallest share the tail overlap cuts in two (tests lower it)
inline int g_rlc_item = 0;                    // kzgb200_dbg_set_tunable("rlc_item", n): run length of the EIP-4844 batch verdict's bucket MSM work items; 0 = default
inline int g_g1fft_split_override = 0;        // kzgb200_dbg_set_tunable("g1fft_split", k): sub-batches (streams) of the staged G1 FFT; 0 = context default
inline int g_pairing_lanes = 0;               // kzgb200_dbg_set_tunable("pairing_lanes", 0 | 8 | 32): see vm_pairing_check (kzgb200_vmsm.cu)
#define KZG_PL32_RESIDENT_WARPS 10            /* pl32 up to 10 checks per SM (of its 12 resident warps); measured crossover with pl8 at ~1 700 checks on 148 SMs (profiles/r02_pairing_sweep.md) */
inline int g_fs_variant = 2;                  // kzgb200_dbg_set_tunable("fiat_shamir", 1 | 2): see launch_fiat_shamir (kzg4844.cuh)
inline int g_fk20_lanes_override = 0;         // kzgb200_dbg_set_tunable("fk20_lanes", L): lanes per 64-point FK20 group for full batches
static inline int msm_variant() {
    if (g_msm_variant_override >= 0) return g_msm_variant_override;
    static const int v = [] {
        const char *e = getenv("KZGB200_MSM_VARIANT");
        int x = e && *e ? atoi(e) : KZG_MSM_DEFAULT_VARIANT;
        return (x < 0 || x > 15) ? KZG_MSM_DEFAULT_VARIANT : x;
    }();
    return v;
}
static inline void launch_msm_fixed(dim3 grid, int tpb, cudaStream_t st, const uint32_t *scalars, const MsmTable &tab, int group_pts, int n_groups, int L,
                                    const int32_t *status, G1 *out) {
    const size_t smem = (size_t)tpb * sizeof(G1) * ((msm_variant() & 8) ? 2 : 1);
#ifdef KZG_MSM_ALL_VARIANTS
    switch (msm_variant()) {
#define KZG_MSM_CASE(V) case V: k_msm_fixed<V><<<grid, tpb, smem, st>>>(scalars, tab, group_pts, n_groups, L, status, out); break;
        KZG_MSM_CASE(0) KZG_MSM_CASE(1) KZG_MSM_CASE(2) KZG_MSM_CASE(3) KZG_MSM_CASE(4) KZG_MSM_CASE(5) KZG_MSM_CASE(6) KZG_MSM_CASE(7)
        KZG_MSM_CASE(8) KZG_MSM_CASE(9) KZG_MSM_CASE(10) KZG_MSM_CASE(11) KZG_MSM_CASE(12) KZG_MSM_CASE(13) KZG_MSM_CASE(14) KZG_MSM_CASE(15)
#undef KZG_MSM_CASE
    }
#else
    k_msm_fixed<KZG_MSM_DEFAULT_VARIANT><<<grid, tpb, smem, st>>>(scalars, tab, group_pts, n_groups, L, status, out);
#endif
}

// out[blob] = sum of the S (<= 32, power of two) partial sums of a blob: one warp per blob, tree through shared memory
static __global__ void __launch_bounds__(32) k_sum_groups(const G1 *__restrict__ partial, int S, const int32_t *__restrict__ status, G1 *__restrict__ out) {
    const int blob = blockIdx.x, t = threadIdx.x;
    if (status && status[blob] != ST_OK) return;
    G1 *sm = reinterpret_cast<G1 *>(msm_smem);
    sm[t] = t < S ? partial[(size_t)blob * S + t] : G1::infinity();
    __syncwarp();
    for (int s = S >> 1; s > 0; s >>= 1) {
        if (t < s) g1_add_ool(&sm[t], &sm[t + s]);
        __syncwarp();
    }
    if (t == 0) out[blob] = sm[0];
}

// ---- table construction (context init) ---------------------------------------------------
// step 1: bases[(j*W + k)] = 2^(bitpos[k]) * P_j  in XYZZ
static __global__ void k_table_bases(const G1Aff *__restrict__ pts, MsmTable tab, G1 *__restrict__ bases) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= tab.npts) return;
    G1 q = G1::from_affine(pts[j]);
    for (int k = 0; k < tab.W; ++k) {
        bases[(size_t)j * tab.W + k] = q;
        if (k + 1 < tab.W) for (int i = 0; i < tab.bits[k]; ++i) q = g1_dbl(q);
    }
}
// XYZZ -> affine, one thread per point (used only at init / for small batches)
static __global__ void k_to_affine(const G1 *__restrict__ in, G1Aff *__restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = g1_to_affine(in[i]);
}
// step 2: one thread per chunk of CH consecutive digits of one (point j, window k): entries
// d = q*CH+1 .. q*CH+CH of d * 2^(bitpos[k]) * P_j, batch-normalised with one inversion per chunk.
#define KZG_TABLE_CHUNK 16
static __global__ void __launch_bounds__(64) k_table_fill(const G1Aff *__restrict__ bases_aff, MsmTable tab) {
    const uint32_t chunks_per_point = tab.row_entries / KZG_TABLE_CHUNK;     // every window has >= 16 entries (c >= 5)
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)tab.npts * chunks_per_point) return;
    size_t j = gid / chunks_per_point;
    uint32_t e0 = (uint32_t)(gid % chunks_per_point) * KZG_TABLE_CHUNK;      // first entry of this chunk inside the row
    int k = 0;
    for (int w = 1; w < tab.W; ++w) if (e0 >= tab.rowoff[w]) k = w;
    unsigned d0 = e0 - tab.rowoff[k];                                        // digits d0+1 .. d0+CH
    G1Aff Q = bases_aff[j * tab.W + k];
    G1Aff *dst = tab.entries + j * tab.row_entries + e0;
    const int cnt = KZG_TABLE_CHUNK;
    if (Q.is_inf()) {
        for (int i = 0; i < cnt; ++i) dst[i] = Q;
        return;
    }
    // start = d0 * Q by MSB-first double-and-add
    G1 cur = G1::infinity();
    for (int bit = 31 - __clz(d0 | 1); bit >= 0; --bit) {
        cur = g1_dbl(cur);
        if ((d0 >> bit) & 1) g1_add_affine(cur, Q);
    }
    G1 pts[KZG_TABLE_CHUNK];
    Fp pre[KZG_TABLE_CHUNK];
    Fp run = Fp::one();
    for (int i = 0; i < cnt; ++i) {
        g1_add_affine(cur, Q);
        pts[i] = cur;
        pre[i] = run;
        run = fp_mul_ni(run, cur.ZZZ);
    }
    Fp inv = fp_inv(run);
    for (int i = cnt - 1; i >= 0; --i) {
        Fp i3 = fp_mul_ni(inv, pre[i]);          // 1/ZZZ_i
        inv = fp_mul_ni(inv, pts[i].ZZZ);
        Fp i2 = fp_mul_ni(fp_sqr_ni(i3), fp_sqr_ni(pts[i].ZZ));
        G1Aff a;
        a.x = fp_mul_ni(pts[i].X, i2);
        a.y = fp_mul_ni(pts[i].Y, i3);
        dst[i] = a;
    }
}

}  // namespace kzg
