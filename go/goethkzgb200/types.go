package goethkzgb200

import "errors"

// Sizes and byte types of the reference (serialization.go:35-95).
const (
	ScalarsPerBlob       = 4096
	SerializedScalarSize = 32
	CompressedG1Size     = 48
	CompressedG2Size     = 96
	ScalarsPerCell       = 64
	BytesPerCell         = ScalarsPerCell * SerializedScalarSize
	CellsPerExtBlob      = 128
)

type (
	Scalar        [SerializedScalarSize]byte
	Blob          [ScalarsPerBlob * SerializedScalarSize]byte
	KZGProof      [CompressedG1Size]byte
	KZGCommitment [CompressedG1Size]byte
	Cell          [BytesPerCell]byte
)

// JSONTrustedSetup has the reference's schema (trusted_setup.go:23-27).
type JSONTrustedSetup struct {
	SetupG2         []string                `json:"g2_monomial"`
	SetupG1Lagrange [ScalarsPerBlob]string `json:"g1_lagrange"`
	SetupG1Monomial [ScalarsPerBlob]string `json:"g1_monomial"`
}

// Sentinel errors: same names as the reference (errors.go:5-22, internal/kzg/errors.go:8) so that callers'
// errors.Is checks keep working, plus the decode errors gnark-crypto returns through serialization.go:108-115.
var (
	ErrVerifyOpeningProof              = errors.New("can not verify opening proof")
	ErrBatchLengthCheck                = errors.New("all designated elements in the batch should have the same size")
	ErrNonCanonicalScalar              = errors.New("scalar is not canonical when interpreted as a big integer in big-endian")
	ErrInvalidCellID                   = errors.New("cell ID should be less than CellsPerExtBlob")
	ErrInvalidRowIndex                 = errors.New("row index should be less than the number of row commitments")
	ErrDeserializeNilInput             = errors.New("cannot not deserialize nil input")
	ErrNumCellIDsNotEqualNumCells      = errors.New("number of cell IDs should be equal to the number of cells")
	ErrCellIDsNotOrdered               = errors.New("cell IDs are not ordered (ascending)")
	ErrFoundInvalidCellID              = errors.New("cell ID should be less than CellsPerExtBlob")
	ErrNotEnoughCellsForReconstruction = errors.New("not enough cells to perform reconstruction")

	ErrInvalidG1Encoding = errors.New("invalid compressed G1 point encoding")
	ErrG1NotOnCurve      = errors.New("compressed G1 point is not on the curve")
	ErrG1NotInSubgroup   = errors.New("G1 point is not in the prime-order subgroup")
	ErrTrustedSetup      = errors.New("malformed trusted setup")
)

// errorFromStatus maps enum kzgb200_status (include/kzgb200.h) to the sentinel errors above.
func errorFromStatus(st int32) error {
	switch st {
	case 0:
		return nil
	case 1:
		return ErrVerifyOpeningProof
	case 2:
		return ErrNonCanonicalScalar
	case 3:
		return ErrInvalidG1Encoding
	case 4:
		return ErrG1NotOnCurve
	case 5:
		return ErrG1NotInSubgroup
	case 6:
		return ErrBatchLengthCheck
	case 7:
		return ErrInvalidCellID
	case 8:
		return ErrCellIDsNotOrdered
	case 9:
		return ErrNotEnoughCellsForReconstruction
	case 10:
		return ErrInvalidRowIndex
	case 12:
		return ErrTrustedSetup
	default:
		return lastError(st)
	}
}
