package goethkzgb200

import (
	"bytes"
	"errors"
	"sync"
	"testing"
)

// Smoke test of the shim on a real GPU (needs libkzgb200.so and a CUDA device): the identities the reference's own
// api_test.go / consensus_specs_test.go rely on, with no YAML dependency.
func TestSmoke(t *testing.T) {
	ctx, err := NewContext4096SecureWithOptions(&Options{CommitWindow: 8, FK20Window: 8})
	if err != nil {
		t.Skipf("no usable GPU context: %v", err)
	}
	var zero Blob
	cm, err := ctx.BlobToKZGCommitment(&zero, 0)
	if err != nil {
		t.Fatal(err)
	}
	inf := KZGCommitment{0xc0} // api.go:47
	if cm != inf {
		t.Fatalf("commitment of the zero blob = %x", cm[:4])
	}
	var blob Blob
	for i := 0; i < ScalarsPerBlob; i++ {
		blob[32*i+31] = byte(i)
		blob[32*i+30] = byte(i >> 8)
	}
	cm, err = ctx.BlobToKZGCommitment(&blob, 0)
	if err != nil {
		t.Fatal(err)
	}
	proof, err := ctx.ComputeBlobKZGProof(&blob, cm, 0)
	if err != nil {
		t.Fatal(err)
	}
	if err := ctx.VerifyBlobKZGProof(&blob, cm, proof); err != nil {
		t.Fatal(err)
	}
	if err := ctx.VerifyBlobKZGProof(&blob, cm, KZGProof(inf)); !errors.Is(err, ErrVerifyOpeningProof) {
		t.Fatalf("wrong proof: %v", err)
	}
	cells, proofs, err := ctx.ComputeCellsAndKZGProofs(&blob, 0)
	if err != nil {
		t.Fatal(err)
	}
	if !bytes.Equal(cells[0][:], blob[:BytesPerCell]) {
		t.Fatal("the first cell is not the head of the blob")
	}
	ids := make([]uint64, 0, 64)
	half := make([]*Cell, 0, 64)
	cms := make([]KZGCommitment, CellsPerExtBlob)
	all := make([]uint64, CellsPerExtBlob)
	for i := 0; i < CellsPerExtBlob; i++ {
		cms[i], all[i] = cm, uint64(i)
		if i%2 == 1 {
			ids, half = append(ids, uint64(i)), append(half, cells[i])
		}
	}
	if err := ctx.VerifyCellKZGProofBatch(cms, all, cells[:], proofs[:]); err != nil {
		t.Fatal(err)
	}
	rc, rp, err := ctx.RecoverCellsAndComputeKZGProofs(ids, half, 0)
	if err != nil {
		t.Fatal(err)
	}
	for i := range rc {
		if *rc[i] != *cells[i] || rp[i] != proofs[i] {
			t.Fatalf("recovered cell / proof %d differs", i)
		}
	}
	// the Context is shared by goroutines, one blob per call (api.go:17-28, verify.go:159-166)
	var wg sync.WaitGroup
	for g := 0; g < 8; g++ {
		wg.Add(1)
		go func() {
			defer wg.Done()
			if c2, err := ctx.BlobToKZGCommitment(&blob, 0); err != nil || c2 != cm {
				t.Errorf("concurrent call: %v", err)
			}
		}()
	}
	wg.Wait()
}
