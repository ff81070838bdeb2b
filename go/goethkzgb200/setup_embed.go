package goethkzgb200

import _ "embed"

// trusted_setup.json is the reference's own file (go-eth-kzg: trusted_setup.json, embedded at
// trusted_setup.go:38-39); copy it next to this file when building.  It is not duplicated in this repository:
// the packed binary form used by the tests lives in go-eth-kzg_b200/data/trusted_setup_4096.bin.
//
//go:embed trusted_setup.json
var embeddedSetupJSON string
