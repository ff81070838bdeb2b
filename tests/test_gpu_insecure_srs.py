"""A context on a SECOND trusted setup: an insecure SRS with a known secret (s = 1234), as the reference's own
structural tests use (internal/kzg_multi/kzg_prove_test.go:18-74: FK20 proofs == naive per-coset quotients;
internal/kzg/srs_test.go:14-36).  With s known every output has a closed form that needs no MSM, FFT or FK20 at all:
    commitment  = [p(s)] G1
    proof_k     = [(p(s) - I_k(s)) / (s^64 - h_k^64)] G1       (I_k = interpolant of cell k on the coset h_k <w_64>)
    blob proof  = [(p(s) - p(z)) / (s - z)] G1
computed here with Python integers; nothing is tied to the ceremony file."""
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu
R = oracle_lib.R_MOD
S = 1234
N = 4096


def _brp(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2)


@pytest.fixture(scope="module")
def setup():
    return oracle_lib.insecure_setup(S)


@pytest.fixture(scope="module")
def ctx(setup):
    import kzgb200
    c = kzgb200.Context(commit_window=8, fk20_window=8, setup=setup)
    yield c
    c.close()


def _eval_at(blob, x):
    """p(x) for the blob's polynomial (evaluations over the bit-reversed 4096th roots, api.go:131), barycentric"""
    w = pow(7, (R - 1) // N, R)
    ev = [int.from_bytes(blob[32 * i:32 * i + 32], "big") for i in range(N)]
    zn = (pow(x, N, R) - 1) * pow(N, -1, R) % R
    acc = 0
    for i in range(N):
        r = pow(w, _brp(i, 12), R)
        acc += ev[i] * r % R * pow((x - r) % R, -1, R)
    return acc % R * zn % R


def test_commitment_and_blob_proof_closed_form(ctx):
    blob = oracle_lib.rand_blob(11 << 20)
    ps = _eval_at(blob, S)
    st, cm = ctx.blob_to_kzg_commitment(blob)
    assert st == 0 and cm == oracle_lib.g1_mul_gen(ps)
    z = 0x1234567890abcdef1234567890abcdef
    st, pf, y = ctx.compute_kzg_proof(blob, z.to_bytes(32, "big"))
    pz = _eval_at(blob, z)
    assert st == 0 and int.from_bytes(y, "big") == pz
    assert pf == oracle_lib.g1_mul_gen((ps - pz) * pow(S - z, -1, R))
    assert ctx.verify_kzg_proof(cm, z.to_bytes(32, "big"), y, pf) == 0
    st, bp = ctx.compute_blob_kzg_proof(blob, cm)
    assert st == 0 and ctx.verify_blob_kzg_proof(blob, cm, bp) == 0
    assert ctx.verify_blob_kzg_proof(blob, cm, pf) == 1


def test_fk20_proofs_equal_naive_quotients(ctx):
    """kzg_prove_test.go:18-74 with the quotient evaluated at the known secret"""
    blob = oracle_lib.rand_blob(12 << 20)
    st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
    assert st == 0
    ps = _eval_at(blob, S)
    w8192 = pow(7, (R - 1) // 8192, R)
    w64 = pow(w8192, 128, R)
    for k in (0, 1, 2, 63, 64, 77, 127):
        h = pow(w8192, _brp(k, 7), R)                       # coset shift of cell k (kzg_multi/srs.go:60-103)
        vals = [int.from_bytes(cells[2048 * k + 32 * j:2048 * k + 32 * j + 32], "big") for j in range(64)]
        # cell k holds p at h * w64^brp6(j); I_k(S) by Lagrange interpolation over that coset
        xs = [h * pow(w64, _brp(j, 6), R) % R for j in range(64)]
        zk = (pow(S, 64, R) - pow(h, 64, R)) % R            # prod (S - x_j) = S^64 - h^64
        ik = 0
        for j in range(64):
            # Z'(x_j) = 64 x_j^63
            ik += vals[j] * zk % R * pow((S - xs[j]) % R * 64 % R * pow(xs[j], 63, R) % R, -1, R)
        ik %= R
        q = (ps - ik) * pow(zk, -1, R) % R
        assert proofs[48 * k:48 * k + 48] == oracle_lib.g1_mul_gen(q), k
    cm = oracle_lib.g1_mul_gen(ps)
    cl = [cells[2048 * i:2048 * i + 2048] for i in range(128)]
    pl = [proofs[48 * i:48 * i + 48] for i in range(128)]
    assert ctx.verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl, pl) == 0
    bad = bytearray(cl[9]); bad[100] ^= 1
    assert ctx.verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl[:9] + [bytes(bad)] + cl[10:], pl) == 1


def test_matches_oracle_on_the_same_insecure_setup(ctx, setup):
    o = oracle_lib.Oracle(setup=setup)
    blob = oracle_lib.rand_blob(13 << 20)
    assert ctx.blob_to_kzg_commitment(blob) == o.blob_to_kzg_commitment(blob)
    st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
    assert (st, cells, proofs) == o.compute_cells_and_kzg_proofs(blob)
    ids = sorted(set(range(0, 128, 3)) | set(range(1, 128, 3)))
    cl = [cells[2048 * i:2048 * i + 2048] for i in ids]
    assert ctx.recover_cells_and_kzg_proofs(ids, cl) == (0, cells, proofs)
