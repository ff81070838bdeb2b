"""A context is shared by several host threads (the reference does the same with goroutines,
verify.go:159-166): concurrent calls must serialise cleanly and give the single-threaded bytes."""
import threading
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu


def test_concurrent_calls_on_one_context():
    import kzgb200
    ctx = kzgb200.Context(commit_window=8, fk20_window=8)
    blobs = [oracle_lib.rand_blob((200 + b) << 20) for b in range(6)]
    ref_c = ctx.blob_to_kzg_commitment_batch(blobs)
    ref_p = ctx.compute_cells_and_kzg_proofs_batch(blobs[:2])
    errs = []

    def worker(k):
        try:
            for _ in range(3):
                if k % 2 == 0:
                    assert ctx.blob_to_kzg_commitment_batch(blobs) == ref_c
                else:
                    assert ctx.compute_cells_and_kzg_proofs_batch(blobs[:2]) == ref_p
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    ctx.close()
    assert not errs, errs
