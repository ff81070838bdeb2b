"""ComputeKZGProof / ComputeBlobKZGProof / RecoverCellsAndComputeKZGProofs through the C ABI."""
import random
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


@pytest.mark.parametrize("fn", ["compute_kzg_proof", "compute_blob_kzg_proof", "recover_cells_and_kzg_proofs"])
def test_spec_vectors(ctx, fn):
    bad = []
    for c in cases(fn):
        got, exp = run_case(ctx, c)
        if got != exp:
            bad.append(c["name"])
    assert not bad, f"{len(bad)}: {bad[:6]}"


def test_blob_proof_batch_matches_oracle(ctx):
    o = oracle_lib.get_oracle()
    blobs = [oracle_lib.rand_blob(b << 20) for b in range(4)]
    cms = [o.blob_to_kzg_commitment(b)[1] for b in blobs]
    got = ctx.compute_blob_kzg_proof_batch(blobs, cms)
    for b, c, (st, p) in zip(blobs, cms, got):
        assert st == 0 and p == o.compute_blob_kzg_proof(b, c)[1]


def test_recover_random_half_equals_compute(ctx):
    """Cfg4: keep a random 64 of 128 cells, recover, compare with the direct computation"""
    blobs = [oracle_lib.rand_blob((10 + b) << 20) for b in range(3)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    ids_list, cells_list = [], []
    for b, (st, cells, proofs) in enumerate(full):
        assert st == 0
        ids = sorted(random.Random(b).sample(range(128), 64 + 5 * b))
        ids_list.append(ids)
        cells_list.append([cells[2048 * i:2048 * (i + 1)] for i in ids])
    got = ctx.recover_cells_and_kzg_proofs_batch(ids_list, cells_list)
    for (st, cells, proofs), (est, ecells, eproofs) in zip(got, full):
        assert st == 0 and cells == ecells and proofs == eproofs


def test_recover_error_statuses_in_batch(ctx):
    blob = oracle_lib.rand_blob(99 << 20)
    st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
    cl = [cells[2048 * i:2048 * (i + 1)] for i in range(128)]
    good = (list(range(0, 128, 2)), [cl[i] for i in range(0, 128, 2)])
    unordered = ([1, 0] + list(range(2, 64)), [cl[1], cl[0]] + cl[2:64])
    few = (list(range(63)), cl[:63])
    bigid = (list(range(63)) + [128], cl[:64])
    got = ctx.recover_cells_and_kzg_proofs_batch([good[0], unordered[0], few[0], bigid[0], good[0]],
                                                 [good[1], unordered[1], few[1], bigid[1], good[1]])
    assert [g[0] for g in got] == [0, 8, 9, 7, 0]
    assert got[0][1] == cells and got[4][2] == proofs
    assert got[1][1] == bytes(262144) and got[2][2] == bytes(6144)


def test_recover_cells_without_proofs(ctx):
    """Context.RecoverCells (api_eip.go:8-15) = the out_proofs == NULL branch of kzgb200_recover_cells_and_kzg_proofs: same cells as
    the full call and as the direct computation, same error statuses, for the 4 valid + the invalid recovery spec vectors"""
    from golden_util import cases, resolve, BadHex
    o = oracle_lib.get_oracle()
    n_valid = 0
    for c in cases("recover_cells_and_kzg_proofs"):
        try:
            ids = list(c["input"]["cell_indices"]); cl = [resolve(x) for x in c["input"]["cells"]]
        except BadHex:
            continue
        st, cells = ctx.recover_cells(ids, cl)
        est = o.recover_cells(ids, cl)
        if c["output"] is None:
            assert st != 0 and est[0] != 0, c["name"]
        else:
            n_valid += 1
            assert st == 0 and cells == b"".join(resolve(x) for x in c["output"][0]) == est[1], c["name"]
    assert n_valid >= 4
    blob = oracle_lib.rand_blob(77 << 20)
    st, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
    ids = sorted(random.Random(7).sample(range(128), 70))
    got = ctx.recover_cells_and_kzg_proofs_batch([ids, list(range(63))], [[cells[2048 * i:2048 * i + 2048] for i in ids], [cells[2048 * i:2048 * i + 2048] for i in range(63)]], proofs=False)
    assert got[0] == (0, cells) and got[1][0] == 9 and got[1][1] == bytes(262144)
