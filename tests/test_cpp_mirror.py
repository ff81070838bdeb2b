"""include/kzgb200.hpp: the C++ mirror of the reference's Context compiles against the C ABI (CPU)
and, on a GPU, reproduces the oracle's bytes through it."""
import os, subprocess, tempfile
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp")
LIBDIR = os.path.join(ROOT, "go-eth-kzg_b200")


def _build(tmp):
    exe = os.path.join(tmp, "host_mirror_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                           "-L", LIBDIR, "-lkzgb200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_cpp_mirror_compiles_and_links():
    with tempfile.TemporaryDirectory() as tmp:
        assert os.path.exists(_build(tmp))


@pytest.mark.gpu
def test_cpp_mirror_matches_oracle():
    import oracle_lib
    o = oracle_lib.get_oracle()
    blob = oracle_lib.rand_blob(77 << 20)
    with tempfile.TemporaryDirectory() as tmp:
        exe = _build(tmp)
        bp, op = os.path.join(tmp, "blob.bin"), os.path.join(tmp, "out.bin")
        open(bp, "wb").write(blob)
        r = subprocess.run([exe, oracle_lib.SETUP, bp, op], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stderr)
        out = open(op, "rb").read()
    cm, pf, cells, cproofs, st = out[:48], out[48:96], out[96:96 + 262144], out[96 + 262144:96 + 262144 + 6144], out[-4:]
    assert cm == o.blob_to_kzg_commitment(blob)[1]
    assert pf == o.compute_blob_kzg_proof(blob, cm)[1]
    _, ecells, eproofs = o.compute_cells_and_kzg_proofs(blob)
    assert cells == ecells and cproofs == eproofs
    assert list(st) == [0, 0, 1, 0]      # verify ok, cell batch ok, corrupted batch rejected, recovery == computation
