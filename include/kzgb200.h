/* kzgb200 -- C ABI of the B200-native KZG engine (libkzgb200.so).
 *
 * This is the drop-in boundary for go-eth-kzg's hot path.  The reference has no FFI of its own
 * (it is pure Go); the boundary it exposes is the method set of *goethkzg.Context.  Each entry
 * point below is what a cgo shim for one of those methods binds; the Go-side stub is shown in
 * INTEGRATION.md.  All citations are relative to the reference repository root.
 *
 * Conventions
 *  - Batched: every call takes n items in flat, caller-owned buffers (blobs 131072 B each,
 *    cells 2048 B, commitments/proofs 48 B, scalars 32 B big-endian -- serialization.go:35-95).
 *  - Buffers may be HOST pointers (pageable or pinned; kzgb200_host_alloc gives pinned memory)
 *    or DEVICE pointers on the context's GPU; the library detects which.  Nothing is retained
 *    after the call returns.  The library works on streams of its own that do NOT synchronise with the
 *    caller's streams: device inputs must be complete (and device outputs unused) when the call is made --
 *    synchronise the producing stream first; outputs are complete when the call returns.
 *  - Return value: KZGB200_OK or a global failure code (bad arguments, CUDA error).  Per-item
 *    outcomes are written to status[i] (int32_t), see enum kzgb200_status.  A per-item failure
 *    leaves that item's outputs zeroed.  The library never aborts and never falls back to the
 *    CPU: without a usable CUDA device kzgb200_ctx_new fails with KZGB200_ERR_CUDA.
 *  - A context is immutable after creation and may be used from several host threads at once
 *    (api.go:17-28 documents the same "create once, share" usage; verify.go:159-166 calls a method from
 *    N goroutines): every GPU of the context has kzgb200_opts.lanes execution lanes, a call takes a free one
 *    (small calls rotate over the GPUs), callers beyond that wait.
 */
#ifndef KZGB200_H
#define KZGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KZGB200_BYTES_PER_BLOB 131072
#define KZGB200_BYTES_PER_CELL 2048
#define KZGB200_CELLS_PER_EXT_BLOB 128
#define KZGB200_BYTES_PER_G1 48
#define KZGB200_BYTES_PER_G2 96
#define KZGB200_BYTES_PER_SCALAR 32

/* Per-item status and global return codes.  0/1 mirror the reference's {nil,
 * kzg.ErrVerifyOpeningProof} (internal/kzg/errors.go:8); the rest mirror errors.go:5-22 and the
 * gnark decode errors surfaced by serialization.go:108-115. */
enum kzgb200_status {
    KZGB200_OK = 0,
    KZGB200_VERIFY_FAILED = 1,          /* kzg.ErrVerifyOpeningProof */
    KZGB200_NON_CANONICAL_SCALAR = 2,   /* ErrNonCanonicalScalar */
    KZGB200_BAD_G1_ENCODING = 3,        /* gnark: invalid flag bits / x >= p / bad infinity */
    KZGB200_NOT_ON_CURVE = 4,           /* gnark: square root does not exist */
    KZGB200_NOT_IN_SUBGROUP = 5,        /* gnark: subgroup check failed */
    KZGB200_LENGTH_MISMATCH = 6,        /* ErrBatchLengthCheck / ErrNumCellIDsNotEqualNumCells */
    KZGB200_BAD_CELL_INDEX = 7,         /* ErrInvalidCellID / ErrFoundInvalidCellID */
    KZGB200_CELL_IDS_NOT_ASCENDING = 8, /* ErrCellIDsNotOrdered */
    KZGB200_NOT_ENOUGH_CELLS = 9,       /* ErrNotEnoughCellsForReconstruction */
    KZGB200_BAD_ROW_INDEX = 10,         /* ErrInvalidRowIndex */
    KZGB200_ERR_ARGS = 11,              /* null pointer / bad size passed to the C ABI */
    KZGB200_ERR_SETUP = 12,             /* trusted setup malformed (api.go:93 ErrMinSRSSize, decode) */
    KZGB200_ERR_CUDA = 100              /* CUDA runtime failure; kzgb200_last_error() has text */
};

typedef struct kzgb200_ctx kzgb200_ctx;

/* Tunables.  Zero-initialise for defaults (one GPU: device 0). */
typedef struct kzgb200_opts {
    int device;          /* CUDA device ordinal, used when n_devices == 0 */
    int commit_window;   /* bits per fixed-base window for the 4096-point Lagrange MSM (7..15); the top 256 mod c windows are one bit wider (MsmTable::plan);
                            table bytes = 4096 * ceil(256/c) * 2^(c-1) * 96.  0 = default */
    int fk20_window;     /* same for the 8192-point FK20 table (twice the bytes).  0 = default */
    int n_devices;       /* 0: the single GPU `device`.  k > 0: the GPUs devices[0..k).  -1: every visible GPU.
                            Each GPU holds a full replica of the tables; a batched call on HOST buffers is cut into
                            contiguous ranges, one per GPU, copied and computed concurrently, results written to disjoint slices of
                            the caller's buffers; the RLC verifiers run one sub-verdict per GPU and AND them (verify.go:152-169 is
                            the reference's own fan-out shape).  Calls on DEVICE buffers run on the GPU that owns the buffers. */
    const int *devices;  /* ordinals when n_devices > 0 (an ordinal may repeat: one more replica on that GPU, for tests) */
    int lanes;           /* concurrent calls per GPU: each lane has its own streams and scratch arena, so `lanes` host threads
                            (goroutines) can be inside the library at once per GPU; further callers wait.  0 = default (2) */
    int reserved[1];
} kzgb200_opts;

/* NewContext4096 (api.go:90-149): g1_monomial / g1_lagrange are 4096 compressed G1 points each
 * in ceremony (natural) order, g2_monomial n_g2 >= 65 compressed G2 points
 * (trusted_setup.go:23-27).  Decompression, the bit-reversal of the Lagrange basis, the FK20
 * table (fk20.go:23-52, toeplitz.go:50-93) and all window tables are built on the GPU. */
int kzgb200_ctx_new(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2_monomial,
                    size_t n_g2, const kzgb200_opts *opts, kzgb200_ctx **out);
void kzgb200_ctx_free(kzgb200_ctx *ctx);

/* Trusted-setup ingestion.
 * kzgb200_parse_trusted_setup_json: the reference's JSONTrustedSetup text (trusted_setup.go:23-27: keys g1_monomial,
 *   g1_lagrange, g2_monomial; hex strings with optional 0x prefix) -> flat 48/96-byte records in ceremony order.
 *   g1 buffers: 4096*48 bytes each; g2_monomial: g2_capacity*96 bytes, *n_g2 receives the count.  Host only.
 * kzgb200_ctx_new_from_json: NewContext4096 (api.go:90-149) on that text.
 * kzgb200_check_trusted_setup: CheckTrustedSetupIsWellFormed (trusted_setup.go:45-83) as batched kernels: every
 *   Lagrange G1, monomial G1 and G2 point must decode and lie in the prime-order subgroup.  *result = OK or the
 *   status of the first failing point in that order; *bad_index (optional) = its index in the concatenation. */
int kzgb200_parse_trusted_setup_json(const char *json, size_t len, uint8_t *g1_monomial, uint8_t *g1_lagrange,
                                     uint8_t *g2_monomial, size_t g2_capacity, size_t *n_g2);
int kzgb200_ctx_new_from_json(const char *json, size_t len, const kzgb200_opts *opts, kzgb200_ctx **out);
int kzgb200_check_trusted_setup(int device, const uint8_t *g1_lagrange, size_t n_lagrange, const uint8_t *g1_monomial, size_t n_monomial,
                                const uint8_t *g2_monomial, size_t n_g2, int32_t *result, size_t *bad_index);
const char *kzgb200_last_error(void);

/* pinned host memory for zero-staging H2D/D2H (portable: usable by every GPU of a multi-GPU context).
 * _interleaved: the pages are spread over all NUMA nodes of the host, for buffers that the GPUs of both sockets read at
 * once (one batch sharded over a multi-GPU context); falls back to the plain allocation where mbind is unavailable. */
void *kzgb200_host_alloc(size_t bytes);
void *kzgb200_host_alloc_interleaved(size_t bytes);
void kzgb200_host_free(void *p);

/* Context.BlobToKZGCommitment (prove.go:13-34), batched. out48: n*48 */
int kzgb200_blob_to_kzg_commitment(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out48, int32_t *status);

/* Context.ComputeBlobKZGProof (prove.go:46-77), batched */
int kzgb200_compute_blob_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48, size_t n,
                                   uint8_t *out48, int32_t *status);

/* Context.ComputeKZGProof (prove.go:85-111), batched: z32 n*32 in, proofs n*48 and y n*32 out */
int kzgb200_compute_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *z32, size_t n,
                              uint8_t *out_proof48, uint8_t *out_y32, int32_t *status);

/* Context.VerifyKZGProof (verify.go:12-41): n independent checks, status[i] in {OK, VERIFY_FAILED, error} */
int kzgb200_verify_kzg_proof(kzgb200_ctx *ctx, const uint8_t *commitments48, const uint8_t *z32, const uint8_t *y32,
                             const uint8_t *proofs48, size_t n, int32_t *status);

/* Context.VerifyBlobKZGProof (verify.go:48-82): n independent checks
 * (also what VerifyBlobKZGProofBatchPar, verify.go:152-169, computes) */
int kzgb200_verify_blob_kzg_proof(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48,
                                  const uint8_t *proofs48, size_t n, int32_t *status);

/* Context.VerifyBlobKZGProofBatch (verify.go:88-145): ONE random-linear-combination verdict for
 * the whole batch in *result (OK / VERIFY_FAILED / first per-item error in index order) */
int kzgb200_verify_blob_kzg_proof_batch(kzgb200_ctx *ctx, const uint8_t *blobs, const uint8_t *commitments48,
                                        const uint8_t *proofs48, size_t n, int32_t *result);

/* Context.ComputeCells (api_eip7594.go:12-26): out_cells n*128*2048 */
int kzgb200_compute_cells(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out_cells, int32_t *status);

/* Context.ComputeCellsAndKZGProofs (api_eip7594.go:28-52): out_proofs n*128*48 */
int kzgb200_compute_cells_and_kzg_proofs(kzgb200_ctx *ctx, const uint8_t *blobs, size_t n, uint8_t *out_cells,
                                         uint8_t *out_proofs, int32_t *status);

/* Context.RecoverCellsAndComputeKZGProofs (api_eip7594.go:144-161), batched over n blobs.
 * cell_ids: flat uint64, counts[i] ids for blob i; cells: the matching flat 2048-byte cells.
 * out_proofs may be NULL (= Context.RecoverCells, api_eip.go:8-15). */
int kzgb200_recover_cells_and_kzg_proofs(kzgb200_ctx *ctx, const uint64_t *cell_ids, const uint64_t *counts,
                                         const uint8_t *cells, size_t n, uint8_t *out_cells, uint8_t *out_proofs,
                                         int32_t *status);

/* Context.VerifyCellKZGProofBatch (api_eip7594.go:163-215).  N cells total, split into n_batches
 * independent RLC verdicts: batch b covers items [batch_offsets[b], batch_offsets[b+1]).
 * commitments48 holds one commitment PER CELL (as the Go API does); de-duplication
 * (api_eip7594.go:238-265) happens inside.  results[b] in {OK, VERIFY_FAILED, error}.
 * A call with many small verdicts (>= 8192 cells in all) first checks ALL its cells as one random linear combination; if
 * that passes every verdict is OK (soundness error < 2^-125, the bound of a verdict's own check), otherwise the verdicts
 * are checked one by one, so results[] never depend on the shortcut.  KZGB200_OPTIMISTIC=0 in the environment disables it. */
int kzgb200_verify_cell_kzg_proof_batch(kzgb200_ctx *ctx, const uint8_t *commitments48, const uint64_t *cell_indices,
                                        const uint8_t *cells, const uint8_t *proofs48, size_t n_cells,
                                        const uint64_t *batch_offsets, size_t n_batches, int32_t *results);

/* Codec validation without a computation attached (the exported deserialisers of serialization.go:108-159, which in
 * the reference return gnark types): status[i] of n compressed G1 points (DeserializeKZGCommitment / DeserializeKZGProof:
 * decode + subgroup check), and of n_items groups of scalars_per_item big-endian scalars (DeserializeScalar: 1,
 * DeserializeBlob: 4096, a cell: 64; NON_CANONICAL_SCALAR if any scalar of the group is >= r). */
int kzgb200_check_g1_points(kzgb200_ctx *ctx, const uint8_t *points48, size_t n, int32_t *status);
int kzgb200_check_scalars(kzgb200_ctx *ctx, const uint8_t *scalars32, size_t n_items, size_t scalars_per_item, int32_t *status);

/* Introspection for benches / DESIGN.md numbers */
typedef struct kzgb200_info {
    int device;
    int sm_count;
    int commit_window, commit_windows_per_scalar;
    int fk20_window, fk20_windows_per_scalar;
    uint64_t commit_table_bytes, fk20_table_bytes;
    double init_ms;
    uint64_t kernel_launches;   /* launches issued by this context since creation (all GPUs, all lanes) */
    int n_devices;              /* GPUs of this context; `device` above is the first */
    int lanes_per_device;
} kzgb200_info;
int kzgb200_get_info(kzgb200_ctx *ctx, kzgb200_info *out);

/* Kernel classes for kzgb200_last_kernel_ms: CUDA-event time per class during the LAST call */
enum kzgb200_kernel_class {
    KZGB200_KC_FR = 0,        /* byte codec + Fr NTT / scalar preparation kernels */
    KZGB200_KC_MSM = 1,       /* k_msm_fixed (fixed-base digit-table MSM) */
    KZGB200_KC_G1FFT = 2,     /* G1 FFT-128 pair of the FK20 pipeline */
    KZGB200_KC_FINALIZE = 3,  /* to-affine + compression */
    KZGB200_KC_VERIFY = 4,    /* verifier glue: coefficient / digit preparation, status merging, final combinations */
    KZGB200_KC_DECODE = 5,    /* k_g1_check: G1 decompression (sqrt) + subgroup check */
    KZGB200_KC_VMSM = 6,      /* variable-base bucket MSMs of the RLC verifiers (k_vmsm_*) */
    KZGB200_KC_PAIRING = 7,   /* k_pairing_lanes */
    KZGB200_N_KERNEL_CLASSES = 8
};
int kzgb200_last_kernel_ms(kzgb200_ctx *ctx, double out[KZGB200_N_KERNEL_CLASSES]);

/* Time of the device-side part of the LAST API call on this context, measured with CUDA events
 * on the lane's stream (kernels only, excluding H2D/D2H when inputs were host buffers); with several GPUs the
 * maximum over the GPUs that took part (kzgb200_last_kernel_ms likewise, per class).  Meaningful when one host
 * thread uses the context (benchmarks); concurrent callers overwrite each other's record. */
double kzgb200_last_device_ms(kzgb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* KZGB200_H */
