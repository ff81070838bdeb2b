package goethkzgb200

import _ "embed"

// The mainnet ceremony output in the packed form the rest of this repository uses
// (go-eth-kzg_b200/data/trusted_setup_4096.bin, byte-identical copy): 4096 x 48 bytes g1_monomial, 4096 x 48 bytes
// g1_lagrange, 65 x 96 bytes g2_monomial, every point in the compressed encoding of the reference's
// trusted_setup.json (trusted_setup.go:23-27, embedded there at :38-39).  Embedding the packed bytes instead of the
// JSON means this package builds from a fresh checkout with no file to copy in.
//
//go:embed trusted_setup_4096.bin
var embeddedSetup []byte

const (
	embeddedG1Bytes = ScalarsPerBlob * CompressedG1Size
	embeddedNumG2   = 65
)
