"""BlobToKZGCommitment through the C ABI on the GPU vs the spec vectors and the oracle."""
import pytest
import oracle_lib
from golden_util import cases
from vector_runner import run_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kzgb200
    c = kzgb200.Context(commit_window=8, fk20_window=8)   # small tables: fast init for tests
    yield c
    c.close()


def test_spec_vectors_blob_to_kzg_commitment(ctx):
    bad = []
    for c in cases("blob_to_kzg_commitment"):
        got, exp = run_case(ctx, c)
        if got != exp:
            bad.append(c["name"])
    assert not bad, bad


def test_random_blobs_match_oracle(ctx):
    o = oracle_lib.get_oracle()
    blobs = [oracle_lib.rand_blob(b << 20) for b in range(6)]
    got = ctx.blob_to_kzg_commitment_batch(blobs)
    for b, (st, cm) in zip(blobs, got):
        assert st == 0
        assert cm == o.blob_to_kzg_commitment(b)[1]


def test_batch_with_bad_blob_in_the_middle(ctx):
    o = oracle_lib.get_oracle()
    good = oracle_lib.rand_blob(123)
    bad = bytearray(good)
    bad[32 * 2111:32 * 2111 + 32] = (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001).to_bytes(32, "big")
    got = ctx.blob_to_kzg_commitment_batch([good, bytes(bad), good])
    exp = o.blob_to_kzg_commitment(good)[1]
    assert got[0] == (0, exp) and got[2] == (0, exp)
    assert got[1][0] == 2 and got[1][1] == bytes(48)


def test_linearity_at_full_window(ctx):
    """size-independent property: commit(a) + commit(b) == commit(a+b) (checked via the oracle's group law)"""
    import ctypes
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    a, b = oracle_lib.rand_blob(1 << 20), oracle_lib.rand_blob(2 << 20)
    s = b"".join(((int.from_bytes(a[i:i + 32], "big") + int.from_bytes(b[i:i + 32], "big")) % R).to_bytes(32, "big")
                 for i in range(0, 131072, 32))
    (sa, ca), (sb, cb), (ss, cs) = ctx.blob_to_kzg_commitment_batch([a, b, s])
    assert sa == sb == ss == 0
    out = ctypes.create_string_buffer(48)
    one = (1).to_bytes(32, "big")
    assert oracle_lib.lib().ko_g1_msm(ca + cb, one + one, ctypes.c_size_t(2), out) == 0
    assert out.raw == cs
