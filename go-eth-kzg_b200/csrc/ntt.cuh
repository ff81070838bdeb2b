// Fr number-theoretic transforms in shared memory (one CTA per transform).
//
// Replaces internal/domain/fft.go:95-144 (fftFrInPlace / FftFr / IfftFr) and
// internal/domain/coset_fft.go:41-70.  The reference always runs decimation-in-frequency and then
// an explicit bit-reversal; here the orderings are chosen so that NO permutation pass exists:
//   * decimation-in-time  takes bit-reversed input  -> natural output
//   * decimation-in-freq. takes natural input       -> bit-reversed output
// and every caller in the EIP-7594 pipeline wants exactly one of those (see fk20.cuh).
//
// Shared-memory layout: limb planes, element i limb k at sm[k*N + i], so that a warp touching 32
// consecutive elements is bank-conflict free for every stage stride >= 32.
//
// Twiddles: one global table roots[t] = w_8192^t (Montgomery), t in [0, 8192); a size-N domain
// uses stride 8192/N, inverse twiddles are roots[(8192 - t) mod 8192].
#pragma once
#include "codec.cuh"

namespace kzg {

static const int ROOTS_N = 8192;

template <int N> __device__ __forceinline__ Fr sm_load(const uint32_t *sm, int i) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.v[k] = sm[k * N + i];
    return r;
}
template <int N> __device__ __forceinline__ void sm_store(uint32_t *sm, int i, const Fr &a) {
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[k * N + i] = a.v[k];
}
__device__ __forceinline__ Fr ld_fr(const Fr *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_fr(Fr *p, const Fr &r) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]);
    q[1] = make_uint4(r.v[4], r.v[5], r.v[6], r.v[7]);
}

// All radix-2 stages of a size-N transform on data already in shared memory.
//   DIT:  input bit-reversed, output natural.   DIF: input natural, output bit-reversed.
//   INVERSE selects w^-1 twiddles (no 1/N scaling here).
template <int LOGN, bool DIT, bool INVERSE>
__device__ __forceinline__ void ntt_smem(uint32_t *sm, const Fr *__restrict__ roots, int tid, int nthreads) {
    constexpr int N = 1 << LOGN;
    constexpr int STRIDE = ROOTS_N / N;
#pragma unroll 1
    for (int s = 0; s < LOGN; ++s) {
        const int log_half = DIT ? s : (LOGN - 1 - s);
        const int half = 1 << log_half;
        for (int b = tid; b < N / 2; b += nthreads) {
            int j = b & (half - 1);
            int i0 = ((b >> log_half) << (log_half + 1)) + j;
            int i1 = i0 + half;
            int t = j * (N >> (log_half + 1)) * STRIDE;          // exponent of w_8192
            if (INVERSE) t = (ROOTS_N - t) & (ROOTS_N - 1);
            Fr x = sm_load<N>(sm, i0), y = sm_load<N>(sm, i1);
            if (DIT) {
                if (t) y = Fr::mul(y, ld_fr(roots + t));
                sm_store<N>(sm, i0, Fr::add(x, y));
                sm_store<N>(sm, i1, Fr::sub(x, y));
            } else {
                Fr d = Fr::sub(x, y);
                if (t) d = Fr::mul(d, ld_fr(roots + t));
                sm_store<N>(sm, i0, Fr::add(x, y));
                sm_store<N>(sm, i1, d);
            }
        }
        __syncthreads();
    }
}

// roots[t] = w_8192^t, one thread per t (square-and-multiply from the generator)
static __global__ void k_init_roots(Fr *roots, Fr gen_mont) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ROOTS_N) return;
    Fr r = Fr::one();
    for (int bit = 12; bit >= 0; --bit) {
        r = Fr::sqr(r);
        if ((t >> bit) & 1) r = Fr::mul(r, gen_mont);
    }
    roots[t] = r;
}

}  // namespace kzg
