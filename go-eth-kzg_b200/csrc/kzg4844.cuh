// EIP-4844 single-point opening path on device: Fiat-Shamir challenge, barycentric evaluation,
// quotient polynomial.  The quotient is then committed with k_msm_fixed (msm.cuh).
//
// Replaces fiatshamir.go:22-40, internal/domain/domain.go:163-235
// (EvaluateLagrangePolynomialWithIndex) and internal/kzg/kzg_prove.go:14-180 (Open,
// computeQuotientPolyOutsideDomain / OnDomain).
#pragma once
#include "ntt.cuh"

namespace kzg {

// ---- SHA-256 (FIPS 180-4), one thread per message -------------------------------------------
__device__ __constant__ const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// one compression; w[16] holds the big-endian message words and is clobbered
__device__ __forceinline__ void sha256_block(uint32_t *h, uint32_t *w) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[i] + w[i & 15];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// computeChallenge (fiatshamir.go:22-33): SHA-256("FSBLOBVERIFY_V1_" || u128_be(4096) || blob || commitment)
// reduced mod r, written as plain little-endian limbs.  One thread per blob.
// message = 32 + 131072 + 48 = 131152 bytes = 2049 full blocks + 16 bytes, then padding.
static __global__ void k_fiat_shamir(const uint8_t *__restrict__ blobs, const uint8_t *__restrict__ commitments, uint32_t *__restrict__ z_out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    uint32_t w[16];
    const uint32_t *bw = reinterpret_cast<const uint32_t *>(blobs + i * 131072);
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(commitments + i * 48);
    // block 0: domain separator (16 B) + 16-byte big-endian 4096 + first 32 bytes of the blob
    w[0] = 0x4653424c; w[1] = 0x4f425645; w[2] = 0x52494659; w[3] = 0x5f56315f;   // "FSBLOBVERIFY_V1_"
    w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 4096;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[8 + k] = __byte_perm(bw[k], 0, 0x0123);
    sha256_block(h, w);
    // blocks 1..2047: blob bytes 32 + 64(b-1) .. ; the blob has 131072-32 = 131040 bytes left = 2047 blocks + 32 bytes.
    // Each thread streams its own blob (lane stride 128 KB), so the next block's 64 bytes are
    // requested one iteration ahead to keep a load in flight under the 64 rounds of compute.
    {
        const uint4 *q = reinterpret_cast<const uint4 *>(bw + 8);
        uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
#pragma unroll 1
        for (int b = 0; b < 2047; ++b) {
            uint4 n0 = v0, n1 = v1, n2 = v2, n3 = v3;
            if (b + 1 < 2047) {
                const uint4 *qn = reinterpret_cast<const uint4 *>(bw + 8 + 16 * (b + 1));
                n0 = __ldg(qn); n1 = __ldg(qn + 1); n2 = __ldg(qn + 2); n3 = __ldg(qn + 3);
            }
            w[0] = __byte_perm(v0.x, 0, 0x0123); w[1] = __byte_perm(v0.y, 0, 0x0123); w[2] = __byte_perm(v0.z, 0, 0x0123); w[3] = __byte_perm(v0.w, 0, 0x0123);
            w[4] = __byte_perm(v1.x, 0, 0x0123); w[5] = __byte_perm(v1.y, 0, 0x0123); w[6] = __byte_perm(v1.z, 0, 0x0123); w[7] = __byte_perm(v1.w, 0, 0x0123);
            w[8] = __byte_perm(v2.x, 0, 0x0123); w[9] = __byte_perm(v2.y, 0, 0x0123); w[10] = __byte_perm(v2.z, 0, 0x0123); w[11] = __byte_perm(v2.w, 0, 0x0123);
            w[12] = __byte_perm(v3.x, 0, 0x0123); w[13] = __byte_perm(v3.y, 0, 0x0123); w[14] = __byte_perm(v3.z, 0, 0x0123); w[15] = __byte_perm(v3.w, 0, 0x0123);
            sha256_block(h, w);
            v0 = n0; v1 = n1; v2 = n2; v3 = n3;
        }
    }
    // block 2048: last 32 blob bytes + first 32 commitment bytes
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = __byte_perm(bw[32768 - 8 + k], 0, 0x0123);
#pragma unroll
    for (int k = 0; k < 8; ++k) w[8 + k] = __byte_perm(cw[k], 0, 0x0123);
    sha256_block(h, w);
    // block 2049: last 16 commitment bytes + 0x80 + zeros + 64-bit length (131152*8 = 1049216 bits)
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = __byte_perm(cw[8 + k], 0, 0x0123);
    w[4] = 0x80000000u;
#pragma unroll
    for (int k = 5; k < 15; ++k) w[k] = 0;
    w[14] = 0; w[15] = 131152u * 8u;
    sha256_block(h, w);
    // digest (big-endian) -> integer mod r (fr.SetBytes reduces): digest < 2^256 < 3r, so <= 2 subtractions
    uint32_t l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = h[7 - k];
    for (int rep = 0; rep < 3; ++rep) {
        if (Fr::geq_limbs(l, FR_MOD)) {
            l[0] = ptx_sub_cc(l[0], FR_MOD[0]);
#pragma unroll
            for (int k = 1; k < 7; ++k) l[k] = ptx_subc_cc(l[k], FR_MOD[k]);
            l[7] = ptx_subc(l[7], FR_MOD[7]);
        }
    }
    uint32_t *o = z_out + i * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = l[k];
}

// The same hash on TWO warps per 32 blobs (round 2).  SHA-256 of one message is a chain of 2050 dependent compressions, so the kernel is
// one thread's latency whatever the batch (3.1 ms).  A compression is 64 rounds whose critical path is ~5 dependent instructions, plus
// the message schedule (48 words of ~7 instructions) which is NOT on that path: warp 1 (producer) fetches block b+1, byte-swaps it,
// expands W[0..63] and stores W[t] + K[t] to shared memory while warp 0 (consumer) runs the rounds of block b from the other buffer.
//   block = 64 threads, grid = ceil(n / 32); shared: 2 buffers x 64 words x 32 lanes = 16 KB.
__device__ __forceinline__ void fs_message_block(uint32_t *w, int b, const uint32_t *bw, const uint32_t *cw) {
    if (b == 0) {
        w[0] = 0x4653424c; w[1] = 0x4f425645; w[2] = 0x52494659; w[3] = 0x5f56315f;   // "FSBLOBVERIFY_V1_"
        w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 4096;
#pragma unroll
        for (int k = 0; k < 8; ++k) w[8 + k] = __byte_perm(bw[k], 0, 0x0123);
    } else if (b <= 2047) {
        const uint4 *q = reinterpret_cast<const uint4 *>(bw + 8 + 16 * (b - 1));
        uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2), v3 = __ldg(q + 3);
        w[0] = __byte_perm(v0.x, 0, 0x0123); w[1] = __byte_perm(v0.y, 0, 0x0123); w[2] = __byte_perm(v0.z, 0, 0x0123); w[3] = __byte_perm(v0.w, 0, 0x0123);
        w[4] = __byte_perm(v1.x, 0, 0x0123); w[5] = __byte_perm(v1.y, 0, 0x0123); w[6] = __byte_perm(v1.z, 0, 0x0123); w[7] = __byte_perm(v1.w, 0, 0x0123);
        w[8] = __byte_perm(v2.x, 0, 0x0123); w[9] = __byte_perm(v2.y, 0, 0x0123); w[10] = __byte_perm(v2.z, 0, 0x0123); w[11] = __byte_perm(v2.w, 0, 0x0123);
        w[12] = __byte_perm(v3.x, 0, 0x0123); w[13] = __byte_perm(v3.y, 0, 0x0123); w[14] = __byte_perm(v3.z, 0, 0x0123); w[15] = __byte_perm(v3.w, 0, 0x0123);
    } else if (b == 2048) {
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __byte_perm(bw[32768 - 8 + k], 0, 0x0123);
#pragma unroll
        for (int k = 0; k < 8; ++k) w[8 + k] = __byte_perm(cw[k], 0, 0x0123);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = __byte_perm(cw[8 + k], 0, 0x0123);
        w[4] = 0x80000000u;
#pragma unroll
        for (int k = 5; k < 15; ++k) w[k] = 0;
        w[15] = 131152u * 8u;
    }
}
#define KZG_FS_BLOCKS 2050
static __global__ void __launch_bounds__(64) k_fiat_shamir2(const uint8_t *__restrict__ blobs, const uint8_t *__restrict__ commitments, uint32_t *__restrict__ z_out, size_t n) {
    __shared__ uint32_t wk[2][64][32];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 32 + lane;
    const size_t ii = i < n ? i : n - 1;                         // idle lanes repeat the last blob (loops stay warp-uniform), nothing is written
    const uint32_t *bw = reinterpret_cast<const uint32_t *>(blobs + ii * 131072);
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(commitments + ii * 48);
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    auto produce = [&](int b) {
        uint32_t w[16];
        fs_message_block(w, b, bw, cw);
        uint32_t (*dst)[32] = wk[b & 1];
#pragma unroll
        for (int t = 0; t < 64; ++t) {
            if (t >= 16) {
                uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
                w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
            }
            dst[t][lane] = w[t & 15] + SHA_K[t];
        }
    };
    if (role == 1) produce(0);
    __syncthreads();
#pragma unroll 1
    for (int b = 0; b < KZG_FS_BLOCKS; ++b) {
        if (role == 1) {
            if (b + 1 < KZG_FS_BLOCKS) produce(b + 1);
        } else {
            const uint32_t (*src)[32] = wk[b & 1];
            uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
            for (int t = 0; t < 64; ++t) {
                uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
                uint32_t ch = (e & f) ^ (~e & g);
                uint32_t t1 = hh + S1 + ch + src[t][lane];
                uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
                uint32_t mj = (a & bb) ^ (a & c) ^ (bb & c);
                uint32_t t2 = S0 + mj;
                hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
            }
            h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        }
        __syncthreads();
    }
    if (role != 0 || i >= n) return;
    // digest (big-endian) -> integer mod r (fr.SetBytes reduces): digest < 2^256 < 3r, so <= 2 subtractions
    uint32_t l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) l[k] = h[7 - k];
    for (int rep = 0; rep < 3; ++rep) {
        if (Fr::geq_limbs(l, FR_MOD)) {
            l[0] = ptx_sub_cc(l[0], FR_MOD[0]);
#pragma unroll
            for (int k = 1; k < 7; ++k) l[k] = ptx_subc_cc(l[k], FR_MOD[k]);
            l[7] = ptx_subc(l[7], FR_MOD[7]);
        }
    }
    uint32_t *o = z_out + i * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = l[k];
}
// launch wrapper: g_fs_variant (tunable "fiat_shamir": 2 = two-warp form (default), 1 = one thread per blob)
static inline void launch_fiat_shamir(cudaStream_t st, const uint8_t *blobs, const uint8_t *commitments, uint32_t *z_out, size_t n) {
    if (!n) return;
    if (g_fs_variant == 1) k_fiat_shamir<<<(unsigned)((n + 31) / 32), 32, 0, st>>>(blobs, commitments, z_out, n);
    else k_fiat_shamir2<<<(unsigned)((n + 31) / 32), 64, 0, st>>>(blobs, commitments, z_out, n);
}

// 32-byte big-endian scalars -> plain limbs with canonical check (DeserializeScalar, serialization.go:153-159)
static __global__ void k_scalars_from_be(const uint8_t *__restrict__ in, uint32_t *__restrict__ out, int32_t *__restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t l[8];
    load_be32(l, in + i * 32);
    if (!fr_is_canonical(l)) atomicMax(&status[i], (int32_t)ST_NON_CANONICAL_SCALAR);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[i * 8 + k] = l[k];
}

// a^-1 in Fr (0 -> 0), binary GCD as fp_inv
static __device__ __noinline__ Fr fr_inv(Fr a) {
    Fr t, r3;
    BinGcd<8>::inverse(t.v, a.v, FR_MOD, FrParams::inv() & 0x7fffffffu);
#pragma unroll
    for (int i = 0; i < 8; ++i) r3.v[i] = FR_R3[i];
    return fr_mul_ni(t, r3);
}

__device__ __forceinline__ Fr fr_to_mont_limbs(const uint32_t *plain) {
    Fr x, r2;
#pragma unroll
    for (int i = 0; i < 8; ++i) { x.v[i] = plain[i]; r2.v[i] = FR_R2[i]; }
    return fr_mul_ni(x, r2);
}

// block-wide inclusive product scan helper over KZG_NTT_THREADS values held in planes sm[8][T]
// (Hillis-Steele; `dir` +1 = prefix, -1 = suffix).  Returns this thread's inclusive product.
__device__ __forceinline__ Fr block_scan_mul(uint32_t *sm, Fr v, int tid, int dir) {
    constexpr int T = KZG_NTT_THREADS;
    sm_store<T>(sm, tid, v);
    __syncthreads();
#pragma unroll 1
    for (int off = 1; off < T; off <<= 1) {
        int src = tid - dir * off;
        Fr o; bool have = src >= 0 && src < T;
        if (have) o = sm_load<T>(sm, src);
        __syncthreads();
        if (have) { v = fr_mul_ni(v, o); sm_store<T>(sm, tid, v); }
        __syncthreads();
    }
    return v;
}

// Open (internal/kzg/kzg_prove.go:14-44), one blob per CTA, in three launches so that the single
// field inversion each blob needs (the batch inversion of the 4096 denominators z - w_i) is done for
// ALL blobs of the chunk at once, one lane per blob, instead of stalling a whole CTA behind one
// thread's 380-product Fermat chain:
//   k_eval_products : den_i = z - w_i; per-thread products; block-wide exclusive prefix x suffix
//                     products -> cex[blob][thread], total[blob]; in-domain index (domain.go:163-171)
//   k_fr_inv_batch  : total[blob] <- 1 / total[blob]
//   k_eval_finish   : per-element inverses, y = (z^n - 1)/n * sum f_i w_i / (z - w_i)
//                     (domain.go:193-235), quotient (kzg_prove.go:81-180)
// The blob's scalars are used in PLAIN form: a Montgomery product of a plain and a Montgomery operand
// is plain, so f_i never needs converting and y / the quotient come out ready to serialise.
//   in : blob bytes (Lagrange evaluations over the bit-reversed domain), z (plain limbs)
//   out: y (32 B big-endian and/or plain limbs, optional), quotient scalars as plain limbs
//        [4096][8] for k_msm_fixed (optional: null = EvaluateLagrangePolynomial only)
// status[blob] must be pre-set (OK or an earlier error); a non-canonical blob sets it in k_eval_finish.
#define KZG_EVAL_PER (4096 / KZG_NTT_THREADS)
__device__ __forceinline__ int eval_root_index(int i) { return (int)(__brev((unsigned)i) >> 20) * 2; }   // bit-reversed 4096-domain: w_4096^brp(i) = w_8192^(2 brp(i))

static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_eval_products(const uint32_t *__restrict__ z_limbs, const Fr *__restrict__ roots,
                                                                   const int32_t *__restrict__ status, Fr *__restrict__ cex,
                                                                   Fr *__restrict__ total, int32_t *__restrict__ dom_index) {
    constexpr int T = KZG_NTT_THREADS, PER = KZG_EVAL_PER;
    __shared__ uint32_t sm[8 * T];
    __shared__ int s_index;
    const int blob = blockIdx.x, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    if (tid == 0) s_index = -1;
    Fr z = fr_to_mont_limbs(z_limbs + (size_t)blob * 8);
    __syncthreads();
    Fr run = Fr::one();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        int i = tid * PER + k;
        Fr den = Fr::sub(z, ld_fr(roots + eval_root_index(i)));
        if (den.is_zero()) { s_index = i; den = Fr::one(); }          // z is in the domain
        run = k ? fr_mul_ni(run, den) : den;
    }
    block_scan_mul(sm, run, tid, +1);
    Fr excl_pre = Fr::one();
    if (tid > 0) excl_pre = sm_load<T>(sm, tid - 1);
    Fr tot = sm_load<T>(sm, T - 1);
    __syncthreads();
    block_scan_mul(sm, run, tid, -1);
    Fr excl_suf = Fr::one();
    if (tid < T - 1) excl_suf = sm_load<T>(sm, tid + 1);
    st_fr(cex + (size_t)blob * T + tid, fr_mul_ni(excl_pre, excl_suf));
    if (tid == 0) { st_fr(total + blob, tot); dom_index[blob] = s_index; }
}

static __global__ void k_fr_inv_batch(Fr *__restrict__ v, const int32_t *__restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || status[i] != ST_OK) return;
    st_fr(v + i, fr_inv(ld_fr(v + i)));
}

static __global__ void __launch_bounds__(KZG_NTT_THREADS) k_eval_finish(const uint8_t *__restrict__ blobs, const uint32_t *__restrict__ z_limbs,
                                                                 const Fr *__restrict__ roots, int32_t *__restrict__ status,
                                                                 const Fr *__restrict__ cex, const Fr *__restrict__ inv_total, const int32_t *__restrict__ dom_index,
                                                                 uint32_t *__restrict__ quotient, uint8_t *__restrict__ y_out,
                                                                 uint32_t *__restrict__ y_limbs, Fr inv_n) {
    constexpr int T = KZG_NTT_THREADS, PER = KZG_EVAL_PER;
    __shared__ uint32_t sm[8 * T];
    __shared__ uint32_t s_y[8];
    const int blob = blockIdx.x, tid = threadIdx.x;
    if (status[blob] != ST_OK) return;
    const int index = dom_index[blob];
    Fr z = fr_to_mont_limbs(z_limbs + (size_t)blob * 8);
    // inv[k] = 1 / den_k: prefix products inside the thread, then the backward sweep from
    // c = 1 / (product of this thread's PER elements) = inv_total * (product of everyone else's)
    Fr inv[PER];
    {
        Fr run = Fr::one();
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            int i = tid * PER + k;
            Fr den = Fr::sub(z, ld_fr(roots + eval_root_index(i)));
            if (i == index) den = Fr::one();
            inv[k] = run;
            run = k ? fr_mul_ni(run, den) : den;
        }
        Fr c = fr_mul_ni(ld_fr(inv_total + blob), ld_fr(cex + (size_t)blob * T + tid));
#pragma unroll
        for (int k = PER - 1; k >= 0; --k) {
            int i = tid * PER + k;
            Fr den = Fr::sub(z, ld_fr(roots + eval_root_index(i)));
            if (i == index) den = Fr::one();
            inv[k] = k ? fr_mul_ni(c, inv[k]) : c;
            if (k) c = fr_mul_ni(c, den);
        }
    }
    // ---- evaluation (f plain, w and inv Montgomery: the products are plain) ----------------------
    const uint8_t *src = blobs + (size_t)blob * 131072;
    int bad = 0;
    Fr acc = Fr::zero();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        int i = tid * PER + k;
        Fr f;
        load_be32(f.v, src + i * 32);
        if (!fr_is_canonical(f.v)) bad = 1;
        if (i == index) {
#pragma unroll
            for (int q = 0; q < 8; ++q) s_y[q] = f.v[q];
        }
        acc = Fr::add(acc, fr_mul_ni(fr_mul_ni(f, ld_fr(roots + eval_root_index(i))), inv[k]));
    }
    bad = __syncthreads_or(bad);
    if (bad) { if (tid == 0) status[blob] = ST_NON_CANONICAL_SCALAR; return; }
    Fr y;                                                              // plain
    if (index >= 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) y.v[q] = s_y[q];
    } else {
        sm_store<T>(sm, tid, acc);
        __syncthreads();
        for (int s = T >> 1; s > 0; s >>= 1) {
            if (tid < s) sm_store<T>(sm, tid, Fr::add(sm_load<T>(sm, tid), sm_load<T>(sm, tid + s)));
            __syncthreads();
        }
        Fr sum = sm_load<T>(sm, 0);
        __syncthreads();
        Fr zn = z;
#pragma unroll 1
        for (int i = 0; i < 12; ++i) zn = fr_sqr_ni(zn);                 // z^4096
        y = fr_mul_ni(fr_mul_ni(Fr::sub(zn, Fr::one()), inv_n), sum);
    }
    if (tid == 0) {
        if (y_out) store_be32(y_out + (size_t)blob * 32, y.v);
        if (y_limbs) {
#pragma unroll
            for (int q8 = 0; q8 < 8; ++q8) y_limbs[(size_t)blob * 8 + q8] = y.v[q8];
        }
    }
    if (!quotient) return;      // evaluation only (verify.go:69, 128)
    // ---- quotient ---------------------------------------------------------------------------------
    // outside: q_i = (f_i - y)/(w_i - z) = -(f_i - y) * inv_i                    (kzg_prove.go:81-111)
    // on domain (index m): same for i != m (den_m was replaced by 1), and
    //   q_m = sum_{i != m} -q_i * w_i / w_m                                       (kzg_prove.go:118-180)
    uint32_t *dst = quotient + (size_t)blob * 4096 * 8;
    Fr qm_acc = Fr::zero();
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        int i = tid * PER + k;
        Fr f;
        load_be32(f.v, src + i * 32);
        Fr q = Fr::neg(fr_mul_ni(Fr::sub(f, y), inv[k]));
        if (index >= 0) {
            if (i == index) q = Fr::zero();                              // overwritten below
            else qm_acc = Fr::sub(qm_acc, fr_mul_ni(q, ld_fr(roots + eval_root_index(i))));
        }
        uint4 *o = reinterpret_cast<uint4 *>(dst + (size_t)i * 8);
        o[0] = make_uint4(q.v[0], q.v[1], q.v[2], q.v[3]);
        o[1] = make_uint4(q.v[4], q.v[5], q.v[6], q.v[7]);
    }
    if (index >= 0) {
        sm_store<T>(sm, tid, qm_acc);
        __syncthreads();
        for (int s = T >> 1; s > 0; s >>= 1) {
            if (tid < s) sm_store<T>(sm, tid, Fr::add(sm_load<T>(sm, tid), sm_load<T>(sm, tid + s)));
            __syncthreads();
        }
        if (tid == index / PER) {
            Fr tot = sm_load<T>(sm, 0);
            // 1 / w_m = w_8192^(8192 - 2 brp(m))
            int t = eval_root_index(index);
            Fr qm = fr_mul_ni(tot, ld_fr(roots + ((ROOTS_N - t) & (ROOTS_N - 1))));
            uint4 *o = reinterpret_cast<uint4 *>(dst + (size_t)index * 8);
            o[0] = make_uint4(qm.v[0], qm.v[1], qm.v[2], qm.v[3]);
            o[1] = make_uint4(qm.v[4], qm.v[5], qm.v[6], qm.v[7]);
        }
    }
}

}  // namespace kzg
