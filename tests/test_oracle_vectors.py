"""Pins the CPU oracle: all 311 consensus-spec vectors + the reference's known-answer tests."""
import hashlib
import pytest
import oracle_lib
from golden_util import cases
from vector_runner import run_case, ALL_FNS


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.get_oracle()


EXPECTED_COUNTS = {"blob_to_kzg_commitment": 11, "compute_kzg_proof": 52, "compute_blob_kzg_proof": 15,
                   "verify_kzg_proof": 122, "verify_blob_kzg_proof": 29, "verify_blob_kzg_proof_batch": 24,
                   "compute_cells_and_kzg_proofs": 11, "recover_cells_and_kzg_proofs": 17,
                   "verify_cell_kzg_proof_batch": 30}


@pytest.mark.parametrize("fn", ALL_FNS)
def test_spec_vectors(oracle, fn):
    cs = cases(fn)
    assert len(cs) == EXPECTED_COUNTS[fn]
    bad = []
    for c in cs:
        got, exp = run_case(oracle, c)
        if got != exp:
            bad.append(c["name"])
    assert not bad, f"{len(bad)}/{len(cs)} mismatches: {bad[:5]}"


def test_fiat_shamir_kat(oracle):
    # fiatshamir_test.go:14-26: zero blob + infinity commitment
    import ctypes
    out = ctypes.create_string_buffer(32)
    oracle.L.ko_compute_challenge(bytes(131072), bytes([0xc0]) + bytes(47), out)
    assert out.raw.hex().startswith("04b7b22a") and out.raw.hex().endswith("6096")


def test_rand_blob_generator():
    # bench_test.go:17-46; GetRandFieldElement(0) = SHA-256(be_int64(0)) mod r = 3b67c9a2...3dfb
    blob = oracle_lib.rand_blob(0)
    r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    for j in (0, 1, 4095):
        d = hashlib.sha256((32 * j).to_bytes(8, "big")).digest()
        assert blob[32 * j:32 * j + 32] == (int.from_bytes(d, "big") % r).to_bytes(32, "big")
    assert blob[:4].hex() == "3b67c9a2"


def test_zero_blob_commitment_is_infinity(oracle):
    # api.go:47 PointAtInfinity
    st, cm = oracle.blob_to_kzg_commitment(bytes(131072))
    assert st == 0 and cm == bytes([0xc0]) + bytes(47)
