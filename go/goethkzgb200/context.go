package goethkzgb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../go-eth-kzg_b200 -lkzgb200 -Wl,-rpath,${SRCDIR}/../../go-eth-kzg_b200
#include <stdlib.h>
#include "kzgb200.h"
*/
import "C"

import (
	"encoding/json"
	"fmt"
	"runtime"
	"unsafe"
)

// Context mirrors goethkzg.Context (api.go:17-28): immutable after creation, safe for concurrent use.  Every GPU of
// the context has Options.Lanes execution lanes (own streams and scratch); a call takes a free lane, single-blob calls
// rotate over the GPUs, so goroutines calling the reference's one-blob methods overlap on the device(s).
type Context struct{ h *C.kzgb200_ctx }

// Options selects the GPUs and table sizes of a context (kzgb200_opts in include/kzgb200.h).  The zero value is one
// GPU (device 0), default windows, two lanes.
type Options struct {
	Devices      []int // CUDA ordinals; nil = device 0; AllDevices = every visible GPU
	Lanes        int   // concurrent calls per GPU, 0 = default
	CommitWindow int   // bits per window of the 4096-point table, 0 = default (13)
	FK20Window   int   // bits per window of the FK20 table, 0 = default (12)
}

// AllDevices as Options.Devices selects every visible GPU: batched calls are sharded over them inside the library.
var AllDevices = []int{-1}

func (o *Options) fill(c *C.kzgb200_opts, pin *[]C.int) {
	if o == nil {
		return
	}
	c.commit_window, c.fk20_window, c.lanes = C.int(o.CommitWindow), C.int(o.FK20Window), C.int(o.Lanes)
	if len(o.Devices) == 1 && o.Devices[0] == -1 {
		c.n_devices = -1
	} else if len(o.Devices) > 0 {
		*pin = make([]C.int, len(o.Devices))
		for i, d := range o.Devices {
			(*pin)[i] = C.int(d)
		}
		c.n_devices = C.int(len(o.Devices))
		c.devices = (*C.int)(C.malloc(C.size_t(len(o.Devices)) * C.size_t(unsafe.Sizeof(C.int(0)))))
		copy(unsafe.Slice(c.devices, len(o.Devices)), *pin)
	}
}

func lastError(st int32) error {
	return fmt.Errorf("kzgb200: status %d: %s", st, C.GoString(C.kzgb200_last_error()))
}

func u8(p unsafe.Pointer) *C.uint8_t { return (*C.uint8_t)(p) }

// NewContext4096 replaces api.go:90-149: the JSON document goes to the library as text; decompression, the
// bit-reversed Lagrange basis, the FK20 table and all window tables are built on the GPU.
func NewContext4096(ts *JSONTrustedSetup) (*Context, error) { return NewContext4096WithOptions(ts, nil) }

// NewContext4096WithOptions is NewContext4096 on chosen GPUs / table sizes.
func NewContext4096WithOptions(ts *JSONTrustedSetup, o *Options) (*Context, error) {
	text, err := json.Marshal(ts)
	if err != nil {
		return nil, err
	}
	var h *C.kzgb200_ctx
	var opts C.kzgb200_opts
	var pin []C.int
	o.fill(&opts, &pin)
	defer C.free(unsafe.Pointer(opts.devices))
	rc := C.kzgb200_ctx_new_from_json((*C.char)(unsafe.Pointer(&text[0])), C.size_t(len(text)), &opts, &h)
	if rc != C.KZGB200_OK {
		return nil, errorFromStatus(int32(rc))
	}
	return wrap(h), nil
}

func wrap(h *C.kzgb200_ctx) *Context {
	c := &Context{h}
	runtime.SetFinalizer(c, func(c *Context) { C.kzgb200_ctx_free(c.h) })
	return c
}

// NewContext4096Secure replaces api.go:53-88 with the embedded mainnet setup (setup_embed.go: the packed bytes of the
// same ceremony output the reference embeds as JSON, trusted_setup.go:38-39).
func NewContext4096Secure() (*Context, error) { return NewContext4096SecureWithOptions(nil) }

// NewContext4096SecureWithOptions is NewContext4096Secure on chosen GPUs / table sizes.
func NewContext4096SecureWithOptions(o *Options) (*Context, error) {
	if len(embeddedSetup) != 2*embeddedG1Bytes+embeddedNumG2*CompressedG2Size {
		return nil, fmt.Errorf("kzgb200: embedded trusted setup has %d bytes", len(embeddedSetup))
	}
	var h *C.kzgb200_ctx
	var opts C.kzgb200_opts
	var pin []C.int
	o.fill(&opts, &pin)
	defer C.free(unsafe.Pointer(opts.devices))
	rc := C.kzgb200_ctx_new(u8(unsafe.Pointer(&embeddedSetup[0])), u8(unsafe.Pointer(&embeddedSetup[embeddedG1Bytes])),
		u8(unsafe.Pointer(&embeddedSetup[2*embeddedG1Bytes])), embeddedNumG2, &opts, &h)
	if rc != C.KZGB200_OK {
		return nil, errorFromStatus(int32(rc))
	}
	return wrap(h), nil
}

// CheckTrustedSetupIsWellFormed replaces trusted_setup.go:45-83 (batched decode + subgroup kernels).
func CheckTrustedSetupIsWellFormed(ts *JSONTrustedSetup) error {
	text, err := json.Marshal(ts)
	if err != nil {
		return err
	}
	g1m := make([]byte, ScalarsPerBlob*CompressedG1Size)
	g1l := make([]byte, ScalarsPerBlob*CompressedG1Size)
	g2 := make([]byte, 4096*CompressedG2Size)
	var nG2 C.size_t
	if rc := C.kzgb200_parse_trusted_setup_json((*C.char)(unsafe.Pointer(&text[0])), C.size_t(len(text)), u8(unsafe.Pointer(&g1m[0])),
		u8(unsafe.Pointer(&g1l[0])), u8(unsafe.Pointer(&g2[0])), 4096, &nG2); rc != C.KZGB200_OK {
		return errorFromStatus(int32(rc))
	}
	var res C.int32_t
	if rc := C.kzgb200_check_trusted_setup(0, u8(unsafe.Pointer(&g1l[0])), ScalarsPerBlob, u8(unsafe.Pointer(&g1m[0])), ScalarsPerBlob,
		u8(unsafe.Pointer(&g2[0])), nG2, &res, nil); rc != C.KZGB200_OK {
		return errorFromStatus(int32(rc))
	}
	return errorFromStatus(int32(res))
}

func both(rc C.int, st C.int32_t) error {
	if rc != C.KZGB200_OK {
		return errorFromStatus(int32(rc))
	}
	return errorFromStatus(int32(st))
}

// ---- EIP-4844 (prove.go, verify.go) ----------------------------------------------------------------------

// BlobToKZGCommitment replaces prove.go:13-34.
func (c *Context) BlobToKZGCommitment(blob *Blob, _ int) (KZGCommitment, error) {
	var out KZGCommitment
	var st C.int32_t
	rc := C.kzgb200_blob_to_kzg_commitment(c.h, u8(unsafe.Pointer(&blob[0])), 1, u8(unsafe.Pointer(&out[0])), &st)
	if err := both(rc, st); err != nil {
		return KZGCommitment{}, err
	}
	return out, nil
}

// ComputeBlobKZGProof replaces prove.go:46-77.
func (c *Context) ComputeBlobKZGProof(blob *Blob, commitment KZGCommitment, _ int) (KZGProof, error) {
	var out KZGProof
	var st C.int32_t
	rc := C.kzgb200_compute_blob_kzg_proof(c.h, u8(unsafe.Pointer(&blob[0])), u8(unsafe.Pointer(&commitment[0])), 1, u8(unsafe.Pointer(&out[0])), &st)
	if err := both(rc, st); err != nil {
		return KZGProof{}, err
	}
	return out, nil
}

// ComputeKZGProof replaces prove.go:85-111.
func (c *Context) ComputeKZGProof(blob *Blob, z Scalar, _ int) (KZGProof, Scalar, error) {
	var proof KZGProof
	var y Scalar
	var st C.int32_t
	rc := C.kzgb200_compute_kzg_proof(c.h, u8(unsafe.Pointer(&blob[0])), u8(unsafe.Pointer(&z[0])), 1, u8(unsafe.Pointer(&proof[0])), u8(unsafe.Pointer(&y[0])), &st)
	if err := both(rc, st); err != nil {
		return KZGProof{}, Scalar{}, err
	}
	return proof, y, nil
}

// VerifyKZGProof replaces verify.go:12-41.
func (c *Context) VerifyKZGProof(commitment KZGCommitment, z, y Scalar, proof KZGProof) error {
	var st C.int32_t
	rc := C.kzgb200_verify_kzg_proof(c.h, u8(unsafe.Pointer(&commitment[0])), u8(unsafe.Pointer(&z[0])), u8(unsafe.Pointer(&y[0])), u8(unsafe.Pointer(&proof[0])), 1, &st)
	return both(rc, st)
}

// VerifyBlobKZGProof replaces verify.go:48-82.
func (c *Context) VerifyBlobKZGProof(blob *Blob, commitment KZGCommitment, proof KZGProof) error {
	var st C.int32_t
	rc := C.kzgb200_verify_blob_kzg_proof(c.h, u8(unsafe.Pointer(&blob[0])), u8(unsafe.Pointer(&commitment[0])), u8(unsafe.Pointer(&proof[0])), 1, &st)
	return both(rc, st)
}

// flattenBlobs copies []*Blob into one contiguous buffer: a slice of Go pointers cannot cross cgo.
func flattenBlobs(blobs []*Blob) ([]byte, error) {
	const blobSize = ScalarsPerBlob * SerializedScalarSize
	flat := make([]byte, len(blobs)*blobSize)
	for i, b := range blobs {
		if b == nil {
			return nil, ErrDeserializeNilInput
		}
		copy(flat[i*blobSize:], b[:])
	}
	return flat, nil
}

func flattenCells(cells []*Cell) ([]byte, error) {
	flat := make([]byte, len(cells)*BytesPerCell)
	for i, cell := range cells {
		if cell == nil {
			return nil, ErrDeserializeNilInput
		}
		copy(flat[i*BytesPerCell:], cell[:])
	}
	return flat, nil
}

func ptrOrNil(b []byte) *C.uint8_t {
	if len(b) == 0 {
		return nil
	}
	return u8(unsafe.Pointer(&b[0]))
}

// commitmentBytes / proofBytes view a slice of 48-byte arrays as bytes without copying.
func commitmentBytes(v []KZGCommitment) []byte {
	if len(v) == 0 {
		return nil
	}
	return unsafe.Slice((*byte)(unsafe.Pointer(&v[0])), len(v)*CompressedG1Size)
}

func proofBytes(v []KZGProof) []byte {
	if len(v) == 0 {
		return nil
	}
	return unsafe.Slice((*byte)(unsafe.Pointer(&v[0])), len(v)*CompressedG1Size)
}

// VerifyBlobKZGProofBatch replaces verify.go:88-145: one random-linear-combination verdict.
func (c *Context) VerifyBlobKZGProofBatch(blobs []*Blob, commitments []KZGCommitment, proofs []KZGProof) error {
	if len(blobs) != len(commitments) || len(blobs) != len(proofs) {
		return ErrBatchLengthCheck
	}
	flat, err := flattenBlobs(blobs)
	if err != nil {
		return err
	}
	var res C.int32_t
	rc := C.kzgb200_verify_blob_kzg_proof_batch(c.h, ptrOrNil(flat), ptrOrNil(commitmentBytes(commitments)), ptrOrNil(proofBytes(proofs)), C.size_t(len(blobs)), &res)
	return both(rc, res)
}

// VerifyBlobKZGProofBatchPar replaces verify.go:152-169: n independent checks, first error in index order.
func (c *Context) VerifyBlobKZGProofBatchPar(blobs []*Blob, commitments []KZGCommitment, proofs []KZGProof) error {
	if len(blobs) != len(commitments) || len(blobs) != len(proofs) {
		return ErrBatchLengthCheck
	}
	if len(blobs) == 0 {
		return nil
	}
	flat, err := flattenBlobs(blobs)
	if err != nil {
		return err
	}
	st := make([]C.int32_t, len(blobs))
	rc := C.kzgb200_verify_blob_kzg_proof(c.h, ptrOrNil(flat), ptrOrNil(commitmentBytes(commitments)), ptrOrNil(proofBytes(proofs)), C.size_t(len(blobs)), &st[0])
	if rc != C.KZGB200_OK {
		return errorFromStatus(int32(rc))
	}
	for _, s := range st {
		if s != 0 {
			return errorFromStatus(int32(s))
		}
	}
	return nil
}

// ---- EIP-7594 (api_eip7594.go, api_eip.go) -------------------------------------------------------------------

func splitCells(flat []byte) [CellsPerExtBlob]*Cell {
	var out [CellsPerExtBlob]*Cell
	for i := range out {
		cell := new(Cell)
		copy(cell[:], flat[i*BytesPerCell:(i+1)*BytesPerCell])
		out[i] = cell
	}
	return out
}

// ComputeCells replaces api_eip7594.go:12-26.
func (c *Context) ComputeCells(blob *Blob, _ int) ([CellsPerExtBlob]*Cell, error) {
	flat := make([]byte, CellsPerExtBlob*BytesPerCell)
	var st C.int32_t
	rc := C.kzgb200_compute_cells(c.h, u8(unsafe.Pointer(&blob[0])), 1, ptrOrNil(flat), &st)
	if err := both(rc, st); err != nil {
		return [CellsPerExtBlob]*Cell{}, err
	}
	return splitCells(flat), nil
}

// ComputeCellsAndKZGProofs replaces api_eip7594.go:28-52.
func (c *Context) ComputeCellsAndKZGProofs(blob *Blob, _ int) ([CellsPerExtBlob]*Cell, [CellsPerExtBlob]KZGProof, error) {
	flat := make([]byte, CellsPerExtBlob*BytesPerCell)
	var proofs [CellsPerExtBlob]KZGProof
	var st C.int32_t
	rc := C.kzgb200_compute_cells_and_kzg_proofs(c.h, u8(unsafe.Pointer(&blob[0])), 1, ptrOrNil(flat), u8(unsafe.Pointer(&proofs[0])), &st)
	if err := both(rc, st); err != nil {
		return [CellsPerExtBlob]*Cell{}, [CellsPerExtBlob]KZGProof{}, err
	}
	return splitCells(flat), proofs, nil
}

func (c *Context) recover(cellIDs []uint64, cells []*Cell, wantProofs bool) ([CellsPerExtBlob]*Cell, [CellsPerExtBlob]KZGProof, error) {
	var noCells [CellsPerExtBlob]*Cell
	var proofs [CellsPerExtBlob]KZGProof
	if len(cellIDs) != len(cells) {
		return noCells, proofs, ErrNumCellIDsNotEqualNumCells // api_eip7594.go:94
	}
	flatIn, err := flattenCells(cells)
	if err != nil {
		return noCells, proofs, err
	}
	flatOut := make([]byte, CellsPerExtBlob*BytesPerCell)
	count := C.uint64_t(len(cellIDs))
	var idPtr *C.uint64_t
	if len(cellIDs) > 0 {
		idPtr = (*C.uint64_t)(unsafe.Pointer(&cellIDs[0]))
	}
	var proofPtr *C.uint8_t
	if wantProofs {
		proofPtr = u8(unsafe.Pointer(&proofs[0]))
	}
	var st C.int32_t
	rc := C.kzgb200_recover_cells_and_kzg_proofs(c.h, idPtr, &count, ptrOrNil(flatIn), 1, ptrOrNil(flatOut), proofPtr, &st)
	if err := both(rc, st); err != nil {
		return noCells, [CellsPerExtBlob]KZGProof{}, err
	}
	return splitCells(flatOut), proofs, nil
}

// RecoverCellsAndComputeKZGProofs replaces api_eip7594.go:144-161.
func (c *Context) RecoverCellsAndComputeKZGProofs(cellIDs []uint64, cells []*Cell, _ int) ([CellsPerExtBlob]*Cell, [CellsPerExtBlob]KZGProof, error) {
	return c.recover(cellIDs, cells, true)
}

// RecoverCells replaces api_eip.go:8-15.
func (c *Context) RecoverCells(cellIDs []uint64, cells []*Cell, _ int) ([CellsPerExtBlob]*Cell, error) {
	out, _, err := c.recover(cellIDs, cells, false)
	return out, err
}

// VerifyCellKZGProofBatch replaces api_eip7594.go:163-215: one verdict over all cells.
func (c *Context) VerifyCellKZGProofBatch(commitments []KZGCommitment, cellIndices []uint64, cells []*Cell, proofs []KZGProof) error {
	n := len(cells)
	if len(commitments) != n || len(cellIndices) != n || len(proofs) != n {
		return ErrBatchLengthCheck // api_eip7594.go:167-171
	}
	if n == 0 {
		return nil // api_eip7594.go:173-175
	}
	flat, err := flattenCells(cells)
	if err != nil {
		return err
	}
	offs := [2]C.uint64_t{0, C.uint64_t(n)}
	var res C.int32_t
	rc := C.kzgb200_verify_cell_kzg_proof_batch(c.h, ptrOrNil(commitmentBytes(commitments)), (*C.uint64_t)(unsafe.Pointer(&cellIndices[0])),
		ptrOrNil(flat), ptrOrNil(proofBytes(proofs)), C.size_t(n), &offs[0], 1, &res)
	return both(rc, res)
}

// ---- additions: the batched entry points the engine is built for --------------------------------------------

// BlobToKZGCommitmentBatch commits to all blobs in one call; errs[i] is the i-th blob's error.
func (c *Context) BlobToKZGCommitmentBatch(blobs []*Blob) ([]KZGCommitment, []error, error) {
	flat, err := flattenBlobs(blobs)
	if err != nil {
		return nil, nil, err
	}
	out := make([]KZGCommitment, len(blobs))
	st := make([]C.int32_t, len(blobs)+1)
	rc := C.kzgb200_blob_to_kzg_commitment(c.h, ptrOrNil(flat), C.size_t(len(blobs)), ptrOrNil(commitmentBytes(out)), &st[0])
	if rc != C.KZGB200_OK {
		return nil, nil, errorFromStatus(int32(rc))
	}
	errs := make([]error, len(blobs))
	for i := range errs {
		errs[i] = errorFromStatus(int32(st[i]))
	}
	return out, errs, nil
}

// ComputeCellsAndKZGProofsBatch returns the flat cells (128*2048 bytes per blob) and proofs (128*48 bytes per blob).
func (c *Context) ComputeCellsAndKZGProofsBatch(blobs []*Blob) (cells, proofs []byte, errs []error, err error) {
	flat, err := flattenBlobs(blobs)
	if err != nil {
		return nil, nil, nil, err
	}
	n := len(blobs)
	cells = make([]byte, n*CellsPerExtBlob*BytesPerCell)
	proofs = make([]byte, n*CellsPerExtBlob*CompressedG1Size)
	st := make([]C.int32_t, n+1)
	rc := C.kzgb200_compute_cells_and_kzg_proofs(c.h, ptrOrNil(flat), C.size_t(n), ptrOrNil(cells), ptrOrNil(proofs), &st[0])
	if rc != C.KZGB200_OK {
		return nil, nil, nil, errorFromStatus(int32(rc))
	}
	errs = make([]error, n)
	for i := range errs {
		errs[i] = errorFromStatus(int32(st[i]))
	}
	return cells, proofs, errs, nil
}

// VerifyCellKZGProofBatches returns one verdict per batch: batch b covers items [offsets[b], offsets[b+1]).
// (The engine checks a call of many small batches as one combined random linear combination first and falls back to the
// per-batch checks only if that does not pass; the results are the same either way.  KZGB200_OPTIMISTIC=0 disables it.)
func (c *Context) VerifyCellKZGProofBatches(commitments []KZGCommitment, cellIndices []uint64, cells []*Cell, proofs []KZGProof, offsets []uint64) ([]error, error) {
	n := len(cells)
	if len(commitments) != n || len(cellIndices) != n || len(proofs) != n || len(offsets) == 0 {
		return nil, ErrBatchLengthCheck
	}
	flat, err := flattenCells(cells)
	if err != nil {
		return nil, err
	}
	nb := len(offsets) - 1
	res := make([]C.int32_t, nb+1)
	var idxPtr *C.uint64_t
	if n > 0 {
		idxPtr = (*C.uint64_t)(unsafe.Pointer(&cellIndices[0]))
	}
	rc := C.kzgb200_verify_cell_kzg_proof_batch(c.h, ptrOrNil(commitmentBytes(commitments)), idxPtr, ptrOrNil(flat), ptrOrNil(proofBytes(proofs)),
		C.size_t(n), (*C.uint64_t)(unsafe.Pointer(&offsets[0])), C.size_t(nb), &res[0])
	if rc != C.KZGB200_OK {
		return nil, errorFromStatus(int32(rc))
	}
	out := make([]error, nb)
	for i := range out {
		out[i] = errorFromStatus(int32(res[i]))
	}
	return out, nil
}
