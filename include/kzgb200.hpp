// kzgb200.hpp -- C++ host-side mirror of go-eth-kzg's `Context` over the C ABI (kzgb200.h).
//
// The reference's boundary is the Go method set of *goethkzg.Context; no Go toolchain exists in
// the build image, so this header is the compiled-language stand-in for the cgo shim: same method
// names, same argument meaning, and the reference's error convention (nil / ErrVerifyOpeningProof /
// any other error) expressed as kzgb200::Error.  Header-only; link against libkzgb200.so.
//
//   Go (reference)                                          here
//   NewContext4096(*JSONTrustedSetup)        api.go:90      Context(g1_monomial, g1_lagrange, g2, n_g2, opts)
//   BlobToKZGCommitment(*Blob, int)          prove.go:13    BlobToKZGCommitment(blob, out)
//   ComputeBlobKZGProof(*Blob, Commitment,…) prove.go:46    ComputeBlobKZGProof(blob, commitment, out)
//   ComputeKZGProof(*Blob, Scalar, int)      prove.go:85    ComputeKZGProof(blob, z, out_proof, out_y)
//   VerifyKZGProof(C, z, y, proof)           verify.go:12   VerifyKZGProof(c, z, y, proof)
//   VerifyBlobKZGProof(*Blob, C, proof)      verify.go:48   VerifyBlobKZGProof(blob, c, proof)
//   VerifyBlobKZGProofBatch([]*Blob, …)      verify.go:88   VerifyBlobKZGProofBatch(blobs, cs, proofs, n)
//   VerifyBlobKZGProofBatchPar(…)            verify.go:152  VerifyBlobKZGProofBatchPar(blobs, cs, proofs, n)
//   ComputeCells(*Blob, int)                 api_eip7594.go:12   ComputeCells(blob, out_cells)
//   ComputeCellsAndKZGProofs(*Blob, int)     api_eip7594.go:28   ComputeCellsAndKZGProofs(blob, out_cells, out_proofs)
//   RecoverCellsAndComputeKZGProofs(ids, cells, int)  :144       RecoverCellsAndComputeKZGProofs(ids, cells, n, out_cells, out_proofs)
//   RecoverCells(ids, cells, int)            api_eip.go:8        RecoverCells(ids, cells, n, out_cells)
//   VerifyCellKZGProofBatch(cs, idx, cells, proofs)   :163       VerifyCellKZGProofBatch(cs, idx, cells, proofs, n)
// numGoRoutines has no meaning on the GPU and is dropped.  Every method also has a *Batch form
// that forwards the caller's flat buffers (host or device pointers) untouched.
#pragma once
#include "kzgb200.h"
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace kzgb200 {

// nil <-> Ok, kzg.ErrVerifyOpeningProof <-> VerifyFailed, everything else is "another error"
enum class Error : int32_t {
    Ok = KZGB200_OK, VerifyFailed = KZGB200_VERIFY_FAILED, NonCanonicalScalar = KZGB200_NON_CANONICAL_SCALAR,
    BadG1Encoding = KZGB200_BAD_G1_ENCODING, NotOnCurve = KZGB200_NOT_ON_CURVE, NotInSubgroup = KZGB200_NOT_IN_SUBGROUP,
    BatchLengthCheck = KZGB200_LENGTH_MISMATCH, InvalidCellID = KZGB200_BAD_CELL_INDEX, CellIDsNotOrdered = KZGB200_CELL_IDS_NOT_ASCENDING,
    NotEnoughCellsForReconstruction = KZGB200_NOT_ENOUGH_CELLS, InvalidRowIndex = KZGB200_BAD_ROW_INDEX, BadArguments = KZGB200_ERR_ARGS,
    BadTrustedSetup = KZGB200_ERR_SETUP, Cuda = KZGB200_ERR_CUDA
};

constexpr size_t BytesPerBlob = KZGB200_BYTES_PER_BLOB, BytesPerCell = KZGB200_BYTES_PER_CELL, CellsPerExtBlob = KZGB200_CELLS_PER_EXT_BLOB;
constexpr size_t CompressedG1Size = KZGB200_BYTES_PER_G1, SerializedScalarSize = KZGB200_BYTES_PER_SCALAR;

class Context {
public:
    // panics (throws) only during creation, like the reference (readme.md:46-50)
    Context(const uint8_t *g1_monomial, const uint8_t *g1_lagrange, const uint8_t *g2_monomial, size_t n_g2, const kzgb200_opts *opts = nullptr) {
        int rc = kzgb200_ctx_new(g1_monomial, g1_lagrange, g2_monomial, n_g2, opts, &h_);
        if (rc != KZGB200_OK) throw std::runtime_error(std::string("kzgb200_ctx_new: ") + kzgb200_last_error());
    }
    // NewContext4096 on the text of a JSONTrustedSetup (trusted_setup.go:23-27, api.go:90)
    explicit Context(const std::string &setup_json, const kzgb200_opts *opts = nullptr) {
        int rc = kzgb200_ctx_new_from_json(setup_json.data(), setup_json.size(), opts, &h_);
        if (rc != KZGB200_OK) throw std::runtime_error(std::string("kzgb200_ctx_new_from_json: ") + kzgb200_last_error());
    }
    ~Context() { kzgb200_ctx_free(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    kzgb200_ctx *handle() const { return h_; }

    // CheckTrustedSetupIsWellFormed (trusted_setup.go:45-83) on flat records: Ok, or the first failing point's error
    static Error CheckTrustedSetupIsWellFormed(const uint8_t *g1_lagrange, size_t n_lagrange, const uint8_t *g1_monomial, size_t n_monomial,
                                               const uint8_t *g2_monomial, size_t n_g2, int device = 0) {
        int32_t res = 0;
        int rc = kzgb200_check_trusted_setup(device, g1_lagrange, n_lagrange, g1_monomial, n_monomial, g2_monomial, n_g2, &res, nullptr);
        return rc != KZGB200_OK ? Error(rc) : Error(res);
    }
    // DeserializeKZGCommitment / DeserializeKZGProof (serialization.go:108-131), validity only
    Error CheckG1Point(const uint8_t *point48) const { int32_t st = 0; int rc = kzgb200_check_g1_points(h_, point48, 1, &st); return both(rc, st); }
    // DeserializeBlob (serialization.go:134-146), validity only
    Error CheckBlob(const uint8_t *blob) const { int32_t st = 0; int rc = kzgb200_check_scalars(h_, blob, 1, 4096, &st); return both(rc, st); }

    // ---- EIP-4844 -------------------------------------------------------------------------------
    Error BlobToKZGCommitment(const uint8_t *blob, uint8_t *out48) const { int32_t st = 0; int rc = kzgb200_blob_to_kzg_commitment(h_, blob, 1, out48, &st); return both(rc, st); }
    Error ComputeBlobKZGProof(const uint8_t *blob, const uint8_t *commitment48, uint8_t *out48) const {
        int32_t st = 0; int rc = kzgb200_compute_blob_kzg_proof(h_, blob, commitment48, 1, out48, &st); return both(rc, st);
    }
    Error ComputeKZGProof(const uint8_t *blob, const uint8_t *z32, uint8_t *out_proof48, uint8_t *out_y32) const {
        int32_t st = 0; int rc = kzgb200_compute_kzg_proof(h_, blob, z32, 1, out_proof48, out_y32, &st); return both(rc, st);
    }
    Error VerifyKZGProof(const uint8_t *c48, const uint8_t *z32, const uint8_t *y32, const uint8_t *proof48) const {
        int32_t st = 0; int rc = kzgb200_verify_kzg_proof(h_, c48, z32, y32, proof48, 1, &st); return both(rc, st);
    }
    Error VerifyBlobKZGProof(const uint8_t *blob, const uint8_t *c48, const uint8_t *proof48) const {
        int32_t st = 0; int rc = kzgb200_verify_blob_kzg_proof(h_, blob, c48, proof48, 1, &st); return both(rc, st);
    }
    // flat buffers of n blobs / commitments / proofs; one random-linear-combination verdict
    Error VerifyBlobKZGProofBatch(const uint8_t *blobs, const uint8_t *cs48, const uint8_t *proofs48, size_t n) const {
        int32_t res = 0; int rc = kzgb200_verify_blob_kzg_proof_batch(h_, blobs, cs48, proofs48, n, &res); return both(rc, res);
    }
    // n independent verifications, first error in index order (errgroup semantics of verify.go:159-168)
    Error VerifyBlobKZGProofBatchPar(const uint8_t *blobs, const uint8_t *cs48, const uint8_t *proofs48, size_t n) const {
        std::vector<int32_t> st(n ? n : 1);
        int rc = kzgb200_verify_blob_kzg_proof(h_, blobs, cs48, proofs48, n, st.data());
        if (rc != KZGB200_OK) return Error(rc);
        for (size_t i = 0; i < n; ++i) if (st[i] != KZGB200_OK) return Error(st[i]);
        return Error::Ok;
    }

    // ---- EIP-7594 -------------------------------------------------------------------------------
    Error ComputeCells(const uint8_t *blob, uint8_t *out_cells /*128*2048*/) const { int32_t st = 0; int rc = kzgb200_compute_cells(h_, blob, 1, out_cells, &st); return both(rc, st); }
    Error ComputeCellsAndKZGProofs(const uint8_t *blob, uint8_t *out_cells, uint8_t *out_proofs /*128*48*/) const {
        int32_t st = 0; int rc = kzgb200_compute_cells_and_kzg_proofs(h_, blob, 1, out_cells, out_proofs, &st); return both(rc, st);
    }
    Error RecoverCellsAndComputeKZGProofs(const uint64_t *cell_ids, const uint8_t *cells, size_t n_cells, uint8_t *out_cells, uint8_t *out_proofs) const {
        uint64_t cnt = n_cells; int32_t st = 0;
        int rc = kzgb200_recover_cells_and_kzg_proofs(h_, cell_ids, &cnt, cells, 1, out_cells, out_proofs, &st);
        return both(rc, st);
    }
    Error RecoverCells(const uint64_t *cell_ids, const uint8_t *cells, size_t n_cells, uint8_t *out_cells) const {
        return RecoverCellsAndComputeKZGProofs(cell_ids, cells, n_cells, out_cells, nullptr);
    }
    // one commitment per cell, as in the Go API; de-duplication happens inside the library
    Error VerifyCellKZGProofBatch(const uint8_t *commitments48, const uint64_t *cell_indices, const uint8_t *cells, const uint8_t *proofs48, size_t n) const {
        uint64_t offs[2] = {0, n}; int32_t res = 0;
        int rc = kzgb200_verify_cell_kzg_proof_batch(h_, commitments48, cell_indices, cells, proofs48, n, offs, 1, &res);
        return both(rc, res);
    }

    // ---- throughput forms: n items per call, per-item status ------------------------------------
    Error BlobToKZGCommitmentBatch(const uint8_t *blobs, size_t n, uint8_t *out48, int32_t *status) const { return Error(kzgb200_blob_to_kzg_commitment(h_, blobs, n, out48, status)); }
    Error ComputeCellsAndKZGProofsBatch(const uint8_t *blobs, size_t n, uint8_t *out_cells, uint8_t *out_proofs, int32_t *status) const {
        return Error(kzgb200_compute_cells_and_kzg_proofs(h_, blobs, n, out_cells, out_proofs, status));
    }

private:
    static Error both(int rc, int32_t st) { return rc != KZGB200_OK ? Error(rc) : Error(st); }
    kzgb200_ctx *h_ = nullptr;
};

}  // namespace kzgb200
