"""The in-library multi-GPU context and its concurrency lanes (include/kzgb200.h: kzgb200_opts.n_devices / devices / lanes;
csrc/kzgb200_api.cu, csrc/shard_plan.hpp).  On a box with one GPU the device list names that GPU twice -- two full
replicas of the (small-window) tables and two lane pools, i.e. exactly the code path of two GPUs: contiguous ranges per
device, one host thread per device, disjoint output slices, per-device sub-verdicts merged on the host.  With >= 2 GPUs
the same tests run across them.  Every result must equal the single-device context's (and the oracle's on a sample)."""
import ctypes, threading
import numpy as np
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    n = torch.cuda.device_count()
    return [0, 1] if n >= 2 else [0, 0]


@pytest.fixture(scope="module")
def ctxs():
    import kzgb200
    one = kzgb200.Context(commit_window=8, fk20_window=8)
    two = kzgb200.Context(commit_window=8, fk20_window=8, devices=_devices(), lanes=2)
    assert two.info()["n_devices"] == 2 and two.info()["lanes_per_device"] == 2
    yield one, two
    one.close(); two.close()


def _cheap_blobs(n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    a[:, :, 0] = 0
    return [a[i].tobytes() for i in range(n)]


def test_sharded_proving_calls_equal_single_device(ctxs):
    one, two = ctxs
    o = oracle_lib.get_oracle()
    blobs = _cheap_blobs(150, 5)                       # >= 2 x 64: commit / proof calls are cut in two
    blobs[70] = bytes([0xff]) * 32 + blobs[70][32:]    # a non-canonical blob in the second range
    c1 = one.blob_to_kzg_commitment_batch(blobs); c2 = two.blob_to_kzg_commitment_batch(blobs)
    assert c1 == c2 and c2[70][0] == 2 and c2[0][0] == 0
    for i in (0, 74, 75, 149):
        assert c2[i] == o.blob_to_kzg_commitment(blobs[i])
    cms = [c for _, c in c2]
    assert one.compute_blob_kzg_proof_batch(blobs, cms) == two.compute_blob_kzg_proof_batch(blobs, cms)
    few = blobs[:40]                                   # >= 2 x 16: cells + proofs are cut in two
    r1 = one.compute_cells_and_kzg_proofs_batch(few); r2 = two.compute_cells_and_kzg_proofs_batch(few)
    assert r1 == r2
    assert r2[39] == o.compute_cells_and_kzg_proofs(few[39])
    # recovery: ragged counts, so the second range starts at a non-trivial offset of the flat id / cell arrays
    ids, cells = [], []
    for b in range(40):
        keep = sorted(np.random.default_rng(b).choice(128, 64 + (b % 5), replace=False).tolist())
        ids.append(keep); cells.append([r2[b][1][2048 * k:2048 * k + 2048] for k in keep])
    ids[33] = ids[33][:60]; cells[33] = cells[33][:60]  # not enough cells
    q1 = one.recover_cells_and_kzg_proofs_batch(ids, cells); q2 = two.recover_cells_and_kzg_proofs_batch(ids, cells)
    assert q1 == q2 and q2[33][0] == 9
    for b in (0, 19, 20, 39):
        assert q2[b] == r2[b]


def test_sharded_verifiers_merge_sub_verdicts(ctxs):
    one, two = ctxs
    blobs = _cheap_blobs(130, 6)
    cms = [c for _, c in one.blob_to_kzg_commitment_batch(blobs)]
    pfs = [p for _, p in one.compute_blob_kzg_proof_batch(blobs, cms)]
    assert two.verify_blob_kzg_proof_batch(blobs, cms, pfs) == 0
    for bad_at in (3, 129):                            # a wrong proof in the first / second device's range
        bad = list(pfs); bad[bad_at] = pfs[(bad_at + 1) % 130]
        assert two.verify_blob_kzg_proof_batch(blobs, cms, bad) == 1
        assert two.verify_blob_kzg_proof_batch_par(blobs, cms, bad) == one.verify_blob_kzg_proof_batch_par(blobs, cms, bad)
    # an undecodable proof on the second device beats a failed pairing on the first (first error, else AND)
    bad = list(pfs); bad[2] = pfs[3]; bad[100] = bytes([0xff]) * 48
    assert two.verify_blob_kzg_proof_batch(blobs, cms, bad) == one.verify_blob_kzg_proof_batch(blobs, cms, bad) == 3
    # cell proofs: 80 blobs x 128 cells = 10240 cells as 80 verdicts (cut by verdict) and as ONE verdict (cut by cell range)
    nb = 80
    full = one.compute_cells_and_kzg_proofs_batch(blobs[:nb])
    cm_flat = b"".join(cms[b] * 128 for b in range(nb)); cells = b"".join(f[1] for f in full); proofs = b"".join(f[2] for f in full)
    idx = np.tile(np.arange(128, dtype=np.uint64), nb)
    N = 128 * nb
    for offs in (np.arange(nb + 1, dtype=np.uint64) * 128, np.array([0, N], dtype=np.uint64)):
        nv = len(offs) - 1
        for corrupt in (None, 5 * 2048 + 40, (N - 3) * 2048 + 7):
            ce = bytearray(cells)
            if corrupt is not None:
                ce[corrupt] ^= 1
            res = [(ctypes.c_int32 * nv)() for _ in range(2)]
            for c, r in zip((one, two), res):
                c._check(c.L.kzgb200_verify_cell_kzg_proof_batch(c.ctx, cm_flat, idx.ctypes.data_as(ctypes.c_void_p), bytes(ce), proofs, ctypes.c_size_t(N),
                                                                  offs.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(nv), r))
            assert list(res[0]) == list(res[1])
            exp = [0] * nv
            if corrupt is not None:
                exp[(corrupt // 2048) // 128 if nv > 1 else 0] = 1
            assert list(res[1]) == exp


def test_batch_offsets_must_cover_all_cells(ctxs):
    one, _ = ctxs
    import kzgb200
    blob = oracle_lib.rand_blob(9 << 20)
    cm = one.blob_to_kzg_commitment(blob)[1]
    _, cells, proofs = one.compute_cells_and_kzg_proofs(blob)
    idx = np.arange(128, dtype=np.uint64)
    res = (ctypes.c_int32 * 2)()
    for offs in ([1, 128], [0, 64], [0, 100, 90], [0, 64, 127]):
        o = np.array(offs, dtype=np.uint64)
        rc = one.L.kzgb200_verify_cell_kzg_proof_batch(one.ctx, cm * 128, idx.ctypes.data_as(ctypes.c_void_p), cells, proofs, ctypes.c_size_t(128),
                                                       o.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(offs) - 1), res)
        assert rc == kzgb200.ERR_ARGS, offs


def test_lanes_serve_concurrent_single_blob_callers(ctxs):
    """N host threads each calling one-blob methods on ONE context (api.go:17-28, verify.go:159-166): results are those of
    sequential calls; with 2 devices x 2 lanes four callers are inside the library at once"""
    _, two = ctxs
    o = oracle_lib.get_oracle()
    blobs = [oracle_lib.rand_blob((40 + i) << 20) for i in range(4)]
    exp_c = [o.blob_to_kzg_commitment(b) for b in blobs]
    exp_p = [o.compute_cells_and_kzg_proofs(b) for b in blobs]
    errs = []

    def worker(i):
        try:
            for rep in range(6):
                k = (i + rep) % 4
                assert two.blob_to_kzg_commitment(blobs[k]) == exp_c[k]
                assert two.compute_cells_and_kzg_proofs(blobs[k]) == exp_p[k]
                cm = exp_c[k][1]
                cl = [exp_p[k][1][2048 * j:2048 * j + 2048] for j in range(128)]
                pl = [exp_p[k][2][48 * j:48 * j + 48] for j in range(128)]
                assert two.verify_cell_kzg_proof_batch([cm] * 128, list(range(128)), cl, pl) == 0
        except Exception as e:          # noqa: BLE001
            errs.append((i, repr(e)))

    th = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_device_buffers_run_on_their_gpu(ctxs):
    """device pointers are never sharded: the call runs on the GPU that owns them and equals the host-buffer result"""
    import torch
    _, two = ctxs
    blobs = _cheap_blobs(140, 7)
    host = two.blob_to_kzg_commitment_batch(blobs)
    dev = "cuda:%d" % _devices()[1]
    d_in = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).to(dev)
    d_out = torch.empty(48 * 140, dtype=torch.uint8, device=dev); d_st = torch.empty(140, dtype=torch.int32, device=dev)
    two.raw_blob_to_kzg_commitment(d_in.data_ptr(), 140, d_out.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    out = bytes(d_out.cpu().numpy())
    assert [(0, out[48 * i:48 * i + 48]) for i in range(140)] == host and int(d_st.abs().sum()) == 0
