"""Size-independent properties at (or near) the BASELINE.json batch sizes, device-pointer and
host-pointer paths of the C ABI, and empty / degenerate batches."""
import ctypes, random
import numpy as np
import pytest
import oracle_lib

pytestmark = pytest.mark.gpu
R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


@pytest.fixture(scope="module")
def ctx():
    import kzgb200
    c = kzgb200.Context(commit_window=10, fk20_window=10)
    yield c
    c.close()


def _cheap_blobs(n, seed=0):
    """n distinct canonical blobs without 4096 SHA calls each: 31 random bytes per scalar (top byte 0)"""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(n, 4096, 32), dtype=np.uint8)
    a[:, :, 0] = 0
    return [a[i].tobytes() for i in range(n)]


def _g1_sum(points):
    out = ctypes.create_string_buffer(48)
    one = (1).to_bytes(32, "big")
    assert oracle_lib.lib().ko_g1_msm(b"".join(points), one * len(points), ctypes.c_size_t(len(points)), out) == 0
    return out.raw


def test_commit_checksum_of_checksums_1024(ctx):
    """sum_b commit(blob_b) == commit(sum_b blob_b): linearity over a 1024-blob batch (several chunks)"""
    n = 1024
    blobs = _cheap_blobs(n, 1)
    got = ctx.blob_to_kzg_commitment_batch(blobs)
    assert all(st == 0 for st, _ in got)
    arr = np.frombuffer(b"".join(blobs), dtype=np.uint8).reshape(n, 4096, 32)
    # scalars are < 2^248, so the column sums of 1024 of them stay < 2^258: reduce mod r in Python ints
    sums = []
    for j in range(4096):
        col = arr[:, j, :].astype(np.uint64)
        acc = 0
        for k in range(32):
            acc = (acc << 8) + int(col[:, k].sum())
        sums.append(acc % R)
    sblob = b"".join(x.to_bytes(32, "big") for x in sums)
    st, cs = ctx.blob_to_kzg_commitment(sblob)
    assert st == 0
    assert _g1_sum([c for _, c in got]) == cs


def test_cells_proofs_roundtrip_verify_and_recover_256(ctx):
    """compute cells+proofs for 256 blobs -> every 128-cell batch verifies; recover from a random half == original"""
    n = 256
    blobs = _cheap_blobs(n, 2)
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    assert all(st == 0 for st, _, _ in full)
    commitments, idx, cells, proofs, offs = [], [], [], [], [0]
    for b, (st, cl, pr) in enumerate(full):
        assert cl[:131072] == blobs[b]
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
        offs.append(len(cells))
    res = ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, offs)
    assert res == [0] * n
    # corrupt one proof in batch 100 (swap two proofs of the same blob)
    bad = list(proofs); bad[100 * 128 + 3], bad[100 * 128 + 4] = bad[100 * 128 + 4], bad[100 * 128 + 3]
    res = ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, offs)
    assert res[100] == 1 and sum(res) == 1
    # recovery from a random 64..100 cells per blob
    ids_list, cells_list = [], []
    for b in range(64):
        ids = sorted(random.Random(b).sample(range(128), 64 + (b % 37)))
        ids_list.append(ids)
        cells_list.append([full[b][1][2048 * i:2048 * (i + 1)] for i in ids])
    rec = ctx.recover_cells_and_kzg_proofs_batch(ids_list, cells_list)
    for b, (st, cl, pr) in enumerate(rec):
        assert st == 0 and cl == full[b][1] and pr == full[b][2]


def test_blob_proof_batch_accept_reject_512(ctx):
    n = 512
    blobs = _cheap_blobs(n, 3)
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    pfs = ctx.compute_blob_kzg_proof_batch(blobs, cms)
    assert all(st == 0 for st, _ in pfs)
    pfs = [p for _, p in pfs]
    assert ctx.verify_blob_kzg_proof_batch(blobs, cms, pfs) == 0
    bad = list(pfs); bad[17] = oracle_lib.load_setup()[0][:48]
    assert ctx.verify_blob_kzg_proof_batch(blobs, cms, bad) == 1
    # idempotence of the prover: same inputs, same bytes
    assert [p for _, p in ctx.compute_blob_kzg_proof_batch(blobs[:8], cms[:8])] == pfs[:8]
    # ComputeKZGProof at z = the Fiat-Shamir challenge must give the same proof
    L = oracle_lib.lib()
    z = ctypes.create_string_buffer(32)
    L.ko_compute_challenge(blobs[5], cms[5], z)
    st, p, y = ctx.compute_kzg_proof(blobs[5], z.raw)
    assert st == 0 and p == pfs[5]
    assert ctx.verify_kzg_proof(cms[5], z.raw, y, p) == 0


def test_empty_batches(ctx):
    assert ctx.blob_to_kzg_commitment_batch([]) == []
    assert ctx.compute_cells_and_kzg_proofs_batch([]) == []
    assert ctx.verify_blob_kzg_proof_batch([], [], []) == 0                     # kzg_verify.go:120-122
    assert ctx.verify_cell_kzg_proof_batch([], [], [], []) == 0                 # api_eip7594.go:173-175
    assert ctx.verify_blob_kzg_proof_batch([bytes(131072)], [], []) == 6       # ErrBatchLengthCheck
    assert ctx.verify_cell_kzg_proof_batch([bytes(48)], [], [], []) == 6


def test_device_pointer_path_matches_host_path(ctx):
    """the same C-ABI call on device pointers (torch tensors) and on host buffers gives identical bytes"""
    torch = pytest.importorskip("torch")
    n = 40
    blobs = _cheap_blobs(n, 4)
    host = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    d_in = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
    d_cells = torch.empty(n * 262144, dtype=torch.uint8, device="cuda")
    d_proofs = torch.empty(n * 6144, dtype=torch.uint8, device="cuda")
    d_st = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.raw_compute_cells_and_kzg_proofs(d_in.data_ptr(), n, d_cells.data_ptr(), d_proofs.data_ptr(), d_st.data_ptr())
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0
    assert bytes(d_cells.cpu().numpy()) == b"".join(c for _, c, _ in host)
    assert bytes(d_proofs.cpu().numpy()) == b"".join(p for _, _, p in host)
    d_cm = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    ctx.raw_blob_to_kzg_commitment(d_in.data_ptr(), n, d_cm.data_ptr(), d_st.data_ptr())
    assert bytes(d_cm.cpu().numpy()) == b"".join(c for _, c in ctx.blob_to_kzg_commitment_batch(blobs))
