"""ComputeCells / ComputeCellsAndKZGProofs through the C ABI on the GPU."""
import pytest
import oracle_lib
from golden_util import cases
from vector_runner import run_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import kzgb200
    c = kzgb200.Context(commit_window=8, fk20_window=8)
    yield c
    c.close()


def test_spec_vectors_compute_cells_and_kzg_proofs(ctx):
    bad = []
    for c in cases("compute_cells_and_kzg_proofs"):
        got, exp = run_case(ctx, c)
        if got != exp:
            bad.append(c["name"])
    assert not bad, bad


def test_compute_cells_agrees_with_cells_and_proofs(ctx):
    # consensus_specs_test.go:389-397: ComputeCells must agree with ComputeCellsAndKZGProofs
    blob = oracle_lib.rand_blob(7 << 20)
    st1, cells1 = ctx.compute_cells(blob)
    st2, cells2, _ = ctx.compute_cells_and_kzg_proofs(blob)
    assert st1 == st2 == 0 and cells1 == cells2
    assert cells1[:131072] == blob      # first 64 cells are the blob itself


def test_random_blobs_match_oracle(ctx):
    o = oracle_lib.get_oracle()
    blobs = [oracle_lib.rand_blob(b << 20) for b in range(3)]
    got = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    for b, (st, cells, proofs) in zip(blobs, got):
        est, ecells, eproofs = o.compute_cells_and_kzg_proofs(b)
        assert st == est == 0
        assert cells == ecells
        assert proofs == eproofs


def test_bad_blob_in_batch(ctx):
    good = oracle_lib.rand_blob(5)
    bad = bytes([0xff]) * 32 + good[32:]
    got = ctx.compute_cells_and_kzg_proofs_batch([good, bad, good])
    assert got[0][0] == 0 and got[2][0] == 0 and got[0][1:] == got[2][1:]
    assert got[1][0] == 2 and got[1][1] == bytes(262144) and got[1][2] == bytes(6144)


def test_small_batch_paths_equal_large_batch_paths(ctx):
    """Small batches take latency-oriented kernels (csrc/g1fft.cuh: the dense one-level G1 transform up to 24 blobs; 64 lanes per
    FK20 group; kzgb200.cu launch_commit_msm: point-range split of the 4096-point MSM below one wave), large ones the staged /
    one-CTA-per-blob forms.  The same blobs must give the same bytes either way, and the oracle's."""
    o = oracle_lib.get_oracle()
    distinct = [oracle_lib.rand_blob((300 + b) << 20) for b in range(5)]
    blobs = [distinct[i % 5] for i in range(40)]
    big = ctx.compute_cells_and_kzg_proofs_batch(blobs)                 # 40 > 24: staged G1 FFT, 8 lanes per group
    for n in (1, 2, 5, 24):                                             # dense transform, 64 / 32 / 16 lanes per group
        small = ctx.compute_cells_and_kzg_proofs_batch(blobs[:n])
        assert small == big[:n], n
    assert big[3] == o.compute_cells_and_kzg_proofs(distinct[3])
    many = [distinct[i % 5] for i in range(500)]                        # 500 >= 3 x 148: one CTA per blob
    cbig = ctx.blob_to_kzg_commitment_batch(many)
    for n in (1, 3, 17, 100, 300):                                      # split factors 32, 32, 32, 8, 2
        assert ctx.blob_to_kzg_commitment_batch(many[:n]) == cbig[:n], n
    assert cbig[4] == o.blob_to_kzg_commitment(distinct[4])


def test_two_level_g1_transform_equals_staged_and_oracle(ctx):
    """Batches of 25 .. 32 blobs take the two-level 16 x 8 G1 transform, 33 .. 64 the 4 x 4 x 4 x 2 one (csrc/g1fft.cuh: k_g1lvl_mul / k_g1lvl_sum).  The same
    blobs through the staged radix-2 form (tunable g1_two_level_max = 0) and through the CPU oracle must give the same bytes; a blob
    with an invalid scalar in the middle of the batch keeps its error status and zeroed outputs."""
    o = oracle_lib.get_oracle()
    distinct = [oracle_lib.rand_blob((500 + b) << 20) for b in range(7)]
    zero = bytes(131072)                                                # constant-zero blob: every MSM sum is the point at infinity
    blobs = [distinct[i % 7] for i in range(45)] + [zero]
    bad = bytes([0xff]) * 32 + distinct[0][32:]
    blobs[17] = bad
    assert ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", 64) == 0  # (the default threshold is 32 blobs)
    try:
        two = ctx.compute_cells_and_kzg_proofs_batch(blobs)             # 46 blobs: two-level
        ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", 0)
        staged = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    finally:
        ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", -1)
    assert two == staged
    # the 4 x 4 x 4 x 2 form (k_g1lvl8_mul) on the same blobs
    assert ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", 0) == 0 and ctx.L.kzgb200_dbg_set_tunable(b"g1_chain4_max", 64) == 0
    try:
        chain4 = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    finally:
        ctx.L.kzgb200_dbg_set_tunable(b"g1_two_level_max", -1); ctx.L.kzgb200_dbg_set_tunable(b"g1_chain4_max", -1)
    assert chain4 == staged
    assert two[17][0] == 2 and two[17][2] == bytes(6144)
    for i in (3, 44, 45):
        assert two[i] == o.compute_cells_and_kzg_proofs(blobs[i]), i
    # default thresholds: 25 and 32 blobs two-level, 33 and 46 the 4 x 4 x 4 x 2 form (sizes around the 32-blob block granularity of the level kernels)
    for n in (25, 32, 33, 46):
        assert ctx.compute_cells_and_kzg_proofs_batch(blobs[:n]) == staged[:n], n
