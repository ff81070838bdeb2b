// Package goethkzgb200 is the Go host side of the B200 KZG engine: a cgo shim that exposes the method set of
// crate-crypto/go-eth-kzg's *Context (api.go, prove.go, verify.go, api_eip7594.go, api_eip.go of the reference)
// on top of the C ABI in include/kzgb200.h (libkzgb200.so).
//
// STATUS: written against the C header, NOT COMPILED OR RUN: the authoring image has no Go toolchain
// (DESIGN.md section 1).  The executable mirrors of the same boundary are include/kzgb200.hpp (C++) and
// go-eth-kzg_b200/kzgb200.py (ctypes), both exercised by the test suite.  Build, once a toolchain exists:
//
//	python go-eth-kzg_b200/build.py                       # libkzgb200.so
//	cd go/goethkzgb200 && go vet . && go test .           # go.mod, the embedded setup and the cgo flags (${SRCDIR}-relative) are in place
//
// The reference's own conformance suite runs against this package by copying consensus_specs_test.go and tests/ from a
// go-eth-kzg checkout next to it and replacing the package clause: the method set, the byte-array types, the sentinel
// errors (errors.Is) and the three-way outcome convention are the same.
//
// Differences from the reference that a caller can observe:
//   - every call runs on the GPU(s) of the context (Options.Devices); there is no CPU fallback, NewContext* fails without a device;
//   - numGoRoutines arguments are accepted and ignored (the reference ignores them on the EIP-7594 paths,
//     api_eip7594.go:54,60);
//   - batch verification draws independent 126-bit coefficients instead of powers of one random scalar
//     (internal/kzg/kzg_verify.go:136-141); verdicts are the same except with probability ~2^-125;
//   - the *Batch methods at the bottom of context.go are additions: one call, thousands of blobs, which is how the
//     engine is meant to be driven (a single-blob call pays launch latency for a 148-SM machine).
package goethkzgb200
