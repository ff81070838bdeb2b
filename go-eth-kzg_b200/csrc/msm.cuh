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
#ifndef KZG_MSM_PREFETCH
#define KZG_MSM_PREFETCH 0
#endif
extern __shared__ unsigned char msm_smem[];
static __global__ void __launch_bounds__(128) k_msm_fixed(const uint32_t *__restrict__ scalars, MsmTable tab, int group_pts,
                                                    int n_groups, int L, const int32_t *__restrict__ status, G1 *__restrict__ out) {
    const int blob = blockIdx.y, t = threadIdx.x;
    if (status && status[blob] != ST_OK) return;
    const int lane = t & (L - 1);
    const int group = blockIdx.x * (blockDim.x / L) + t / L;
    const int total = group_pts * n_groups;
    G1 acc = G1::infinity();
    if (group < n_groups) {
        const uint32_t *sc = scalars + ((size_t)blob * total + (size_t)group * group_pts) * 8;
        for (int j = lane; j < group_pts; j += L) {
            DigitStream ds;
            {
                const uint4 *q = reinterpret_cast<const uint4 *>(sc + (size_t)j * 8);
                uint4 a = __ldg(q), b = __ldg(q + 1);
                uint32_t l[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                ds.init(l);
            }
            const G1Aff *row = tab.entries + (size_t)(group * group_pts + j) * tab.row_entries;
#if KZG_MSM_PREFETCH
            // software pipeline: the gather of window k + 1 (a random 96-byte read of a multi-GB table: DRAM latency) is
            // issued before the ten products of window k's addition, so it lands underneath them
            int d = ds.next(tab.bitpos[0], tab.bits[0]);
            G1Aff e;
            if (d) e = load_aff(row + tab.rowoff[0] + ((d < 0 ? -d : d) - 1));
            for (int k = 0; k < tab.W; ++k) {
                int dn = 0;
                G1Aff en;
                if (k + 1 < tab.W) {
                    dn = ds.next(tab.bitpos[k + 1], tab.bits[k + 1]);
                    if (dn) en = load_aff(row + tab.rowoff[k + 1] + ((dn < 0 ? -dn : dn) - 1));
                }
                if (d) {
                    if (d < 0) e.y = Fp::neg(e.y);
                    g1_add_affine<MulInline>(acc, e);
                }
                d = dn; e = en;
            }
#else
            for (int k = 0; k < tab.W; ++k) {
                int d = ds.next(tab.bitpos[k], tab.bits[k]);
                if (d == 0) continue;
                int mag = d < 0 ? -d : d;
                G1Aff e = load_aff(row + tab.rowoff[k] + (mag - 1));
                if (d < 0) e.y = Fp::neg(e.y);
                g1_add_affine<MulInline>(acc, e);
            }
#endif
        }
    }
    // tree reduction over the L lanes of each group through shared memory
    G1 *sm = reinterpret_cast<G1 *>(msm_smem);
    sm[t] = acc;
    __syncthreads();
    for (int s = L >> 1; s > 0; s >>= 1) {
        if (lane < s) g1_add_ool(&sm[t], &sm[t + s]);
        __syncthreads();
    }
    if (lane == 0 && group < n_groups) out[(size_t)blob * n_groups + group] = sm[t];
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
