/* kzgb200 debug / unit-test hooks.  NOT part of the drop-in boundary: they exist so that
 * tests/ can check the device field and group arithmetic against Python integers and the oracle
 * one primitive at a time, and so that bench.py can measure the IMAD roofline denominator.
 * Operands are plain (non-Montgomery) little-endian 32-bit limbs: 12 per Fp, 8 per Fr. */
#ifndef KZGB200_DEBUG_H
#define KZGB200_DEBUG_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* op: 0 mul, 1 add, 2 sub, 3 inverse of a (Fp only), 4 square of a */
int kzgb200_dbg_fp_op(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op);
int kzgb200_dbg_fr_op(const uint32_t *a, const uint32_t *b, uint32_t *out, int n, int op);
/* compressed points in/out. op: 0 a + b (XYZZ + XYZZ with non-trivial ZZ), 1 a + b (mixed), 2 2a */
int kzgb200_dbg_g1_op(const uint8_t *a48, const uint8_t *b48, uint8_t *out48, int n, int op);
/* out = [scalar] point (generic 255-bit scalar multiplication kernel) */
int kzgb200_dbg_g1_mul(const uint8_t *p48, const uint8_t *s32, uint8_t *out48, int n);
/* out[i] = 1 iff e(a_i, Q[qa_i]) * e(b_i, Q[qb_i]) == 1, Q = {G2, [s]G2, [s^64]G2} of the context */
struct kzgb200_ctx;
int kzgb200_dbg_pairing(struct kzgb200_ctx *ctx, const uint8_t *a48, const int *qa, const uint8_t *b48, const int *qb, int *out, int n);
/* gamma[1], lines[0].A[0], lines[0].B[0], lines[0].A[67] as 4 x (c0,c1) x 12 plain limbs */
int kzgb200_dbg_dump_pairing(struct kzgb200_ctx *ctx, uint32_t *out96);
/* GLV split + signed base-16 recoding used by the verifiers' bucket MSMs: for n plain scalars s < r (8 limbs each)
 * writes 64 digits per scalar, digits[0..31] of k1 and [32..63] of k2 with s = k1 - k2 x^2 (mod r), each in [-8, 8] */
int kzgb200_dbg_glv_digits(const uint32_t *s, int8_t *digits, int n);
/* the verifiers' variable-base bucket MSM on its own: out48 = sum_i [s_i] P_i for n compressed points and n plain
 * 255-bit scalars (one verdict, work items of 128 points), through k_vmsm_buckets / reduce / combine */
int kzgb200_dbg_vmsm(struct kzgb200_ctx *ctx, const uint8_t *p48, const uint32_t *s, int n, uint8_t *out48);
/* self-consistency of the device G2 code on one compressed point: bit 0 decodes, 1 on the twist, 2 2Q on the twist,
 * 3 3Q on the twist, 4 2(2Q) == 3Q + Q, 5 Q + Q (addition's doubling branch) == 2Q, 6 Q + (-Q) == O,
 * 7 g2_in_subgroup(Q), 8 [r]Q == O by a ladder written out in the test kernel (7 and 8 must agree) */
int kzgb200_dbg_g2_selftest(const uint8_t *in96, int *mask);
/* host bookkeeping of VerifyCellKZGProofBatch (csrc/cell_plan.hpp; no GPU involved): per-verdict de-duplication of the
 * commitments, grouping of cells by unique commitment, range check of the cell indices and the kernels' work-item
 * lists, as JSON text in a thread-local buffer; NULL if batch_offsets is not monotone / out of range */
const char *kzgb200_dbg_plan_cell_batches_json(const uint8_t *commitments48, const uint64_t *cell_indices, size_t n_cells,
                                               const uint64_t *batch_offsets, size_t n_batches, uint64_t item, uint64_t large, uint64_t row_item);
/* the same with the run length of the large verdicts' column work items given separately (0 = item) */
const char *kzgb200_dbg_plan_cell_batches_json_l(const uint8_t *commitments48, const uint64_t *cell_indices, size_t n_cells,
                                                 const uint64_t *batch_offsets, size_t n_batches, uint64_t item, uint64_t large, uint64_t row_item, uint64_t l_item);
/* host-side sharding of a batched call over the GPUs of a context (csrc/shard_plan.hpp; no GPU involved), as JSON text in a
 * thread-local buffer: "units" = contiguous ranges of n_units independent units over n_dev devices with at least min_per_dev
 * each; "verdicts" = ranges of verdicts balanced by cell count (batch_offsets may be NULL); "merged" = the three-way merge of
 * the sub-verdicts of one logical verdict (first error in index order, else VERIFY_FAILED if any, else OK) */
const char *kzgb200_dbg_shard_plan_json(size_t n_units, size_t n_dev, size_t min_per_dev, const uint64_t *batch_offsets, size_t n_batches,
                                        size_t min_cells_per_dev, const int32_t *sub_verdicts, size_t n_sub);
/* aggregate host -> device GB/s of n GPUs copying their slices of ONE pinned host buffer at the same time (plain or
 * NUMA-interleaved pinned memory): the ceiling of the end-to-end numbers of a multi-GPU context */
int kzgb200_dbg_h2d_bandwidth(const int *devices, int n, size_t bytes_per_dev, int interleaved, double *gbps);
/* host-only unit-test hook: source index and twiddle exponent (of w_128) of term j of output o at `level` of the multi-level forms of the
 * FK20 G1 transform (csrc/g1fft.cuh; form 0: 128 = 16 x 8, levels 0..3; form 1: 128 = 4 x 4 x 4 x 2, levels 0..7) */
int kzgb200_dbg_g1_level_term(int form, int level, int o, int j, int *src, int *e);
/* experiments: run-time tunables of the proving paths (every setting is bit-exact).
 *   "msm_variant": k_msm_fixed variant, bit 0: out-of-line field products, bit 1: cp.async-staged gather, bit 2: one-reduction
 *                  Y3, bit 3: accumulator in shared memory + 4 CTAs/SM; -1 = default (or the KZGB200_MSM_VARIANT environment variable)
 *   "fk20_lanes":  lanes per 64-point FK20 group for full batches (4, 8 or 16; 0 = default)
 *   "g1fft_split": sub-batches (streams) of the staged G1 FFT, 0 = context default
 *   "g1fft_dual": 1 = the G1 FFT stage kernel runs every pair of independent field products of a doubling / addition inside one out-of-line
 *                  body at 2 CTAs per SM (default 0: measured slower); "decode_dual": 1 (default) = paired squarings in the subgroup test
 *   "pairing_lanes": lanes per pairing check: 32 (one check per warp, latency form), 8 (throughput form), 0 = by batch size (default)
 *   "fiat_shamir": SHA-256 of the Fiat-Shamir challenge: 2 = two warps per 32 blobs (message schedule on a producer warp; default), 1 = one thread per blob
 *   "vmsm_policy": field products of the verifiers' bucket accumulation: 0 inlined, 1 out-of-line + one-reduction Y3 (default),
 *                  2 inlined + one-reduction Y3, 3 out-of-line
 *   "optimistic": 1 (default) = VerifyCellKZGProofBatch checks a call of many small verdicts as one combined verdict first, 0 = per-verdict checks only
 *   "large_window": 4 (default) | 8 = window bits of the column MSMs of large (>= 4096-cell) verdicts; "large_item": run length of their work items (0 = default)
 *   "verify_overlap": 1 (default) = the cell verifier's interpolation chain runs on a side stream beside the proofs' decode, 0 = one stream
 *   "g1_two_level_max": chunks of up to this many blobs take the two-level (16 x 8) G1 transform instead of the staged one (-1 = default 32, 0 = never)
 *   "g1_chain4_max": chunks above g1_two_level_max and up to this many blobs take the 4 x 4 x 4 x 2 G1 transform (-1 = default 64, 0 = never)
 *   "proof_pieces": 2..4 (default 3) = pieces the host-buffer ComputeKZGProof / ComputeBlobKZGProof paths cut their input into
 *   "tail_split": 1 = a share of >= "tail_min_cells" cells of VerifyCellKZGProofBatch is cut 7/8 + 1/8 onto two lanes of its GPU, the small part gated
 *                  behind the big part's decode so that it runs during the big part's latency-bound tails; 0 (default: measured, no gain) = one lane per share
 *   "tail_min_cells": smallest share (in cells) that "tail_split" cuts in two (default 131072)
 *   "rlc_item": run length of the EIP-4844 batch verdict's bucket-MSM work items (0 = default 128) */
int kzgb200_dbg_set_tunable(const char *name, int v);
/* dependency-free integer multiply-add microbenchmark: device-wide instructions*lanes per second.
 * mode 0: mad.lo.u32 (IMAD), 1: mad.hi.u32 (IMAD.HI), 2: mad.wide.u32 (IMAD.WIDE, 32x32+64);
 * mode 3 / 4: a bare dependent chain of the library's Fp::mul / Fp::sqr per thread -> field operations per second */
int kzgb200_bench_imad(int device, int mode, double *per_s, double *ms_out);
#ifdef __cplusplus
}
#endif
#endif
