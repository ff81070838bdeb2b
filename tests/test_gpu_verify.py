"""The verifiers through the C ABI on the GPU: spec vectors (three-way outcome) + synthetic batches."""
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


@pytest.mark.parametrize("fn", ["verify_kzg_proof", "verify_blob_kzg_proof", "verify_blob_kzg_proof_batch", "verify_cell_kzg_proof_batch"])
def test_spec_vectors(ctx, fn):
    bad = []
    for c in cases(fn):
        got, exp = run_case(ctx, c)
        if got != exp:
            bad.append((c["name"], got, exp))
    assert not bad, f"{len(bad)}: {bad[:6]}"


def test_batch_par_agrees_with_batch_on_vectors(ctx):
    # consensus_specs_test.go:342-344
    from golden_util import resolve, BadHex
    for c in cases("verify_blob_kzg_proof_batch"):
        try:
            blobs = [resolve(x) for x in c["input"]["blobs"]]
            cms = [resolve(x) for x in c["input"]["commitments"]]
            pfs = [resolve(x) for x in c["input"]["proofs"]]
        except BadHex:
            continue
        if any(len(b) != 131072 for b in blobs) or any(len(x) != 48 for x in cms + pfs):
            continue
        a = ctx.verify_blob_kzg_proof_batch(blobs, cms, pfs)
        b = ctx.verify_blob_kzg_proof_batch_par(blobs, cms, pfs)
        norm = lambda s: True if s == 0 else False if s == 1 else None
        assert norm(a) == norm(b), c["name"]


def test_synthetic_blob_batch_accept_and_reject(ctx):
    """Cfg2 in miniature: commit, prove, batch-verify (accept); corrupt one proof (reject)"""
    blobs = [oracle_lib.rand_blob((40 + b) << 20) for b in range(9)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    pfs = [p for _, p in ctx.compute_blob_kzg_proof_batch(blobs, cms)]
    assert ctx.verify_blob_kzg_proof_batch(blobs, cms, pfs) == 0
    assert ctx.verify_blob_kzg_proof_each(blobs, cms, pfs) == [0] * 9
    gen = oracle_lib.load_setup()[0][:48]          # G1 generator: a valid point, wrong proof
    bad = list(pfs); bad[5] = gen
    assert ctx.verify_blob_kzg_proof_batch(blobs, cms, bad) == 1
    each = ctx.verify_blob_kzg_proof_each(blobs, cms, bad)
    assert each[5] == 1 and sum(each) == 1
    # oracle agrees
    o = oracle_lib.get_oracle()
    assert o.verify_blob_kzg_proof_batch(blobs, cms, pfs) == 0 and o.verify_blob_kzg_proof_batch(blobs, cms, bad) == 1


def test_synthetic_cell_batches(ctx):
    """Cfg5 in miniature: several independent 128-cell batches in one call, one corrupted"""
    blobs = [oracle_lib.rand_blob((60 + b) << 20) for b in range(3)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs, offs = [], [], [], [], [0]
    for b, (st, cl, pr) in enumerate(full):
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
        offs.append(len(cells))
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, offs) == [0, 0, 0]
    # one batch spanning all three blobs (multi-commitment rows)
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, [0, len(cells)]) == [0]
    # corrupt one cell of the middle batch (swap with a different cell of the same blob)
    bad = list(cells); bad[128 + 7] = cells[128 + 8]
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, bad, proofs, offs) == [0, 1, 0]
    # out-of-range cell index -> error for that batch only
    bidx = list(idx); bidx[2 * 128 + 3] = 128
    assert ctx.verify_cell_kzg_proof_batches(commitments, bidx, cells, proofs, offs) == [0, 0, 7]
    # a subset with repeated cells and an empty batch
    sub = [5, 5, 9, 127, 0, 5]
    r = ctx.verify_cell_kzg_proof_batches([commitments[i] for i in sub], [idx[i] for i in sub], [cells[i] for i in sub],
                                          [proofs[i] for i in sub], [0, 0, len(sub)])
    assert r == [0, 0]
    # nothing at all: one empty verdict (api_eip7594.go:173-175: an empty batch verifies), and no verdicts
    assert ctx.verify_cell_kzg_proof_batches([], [], [], [], [0, 0]) == [0]
    assert ctx.verify_cell_kzg_proof_batches([], [], [], [], [0]) == []


def test_cell_batch_larger_than_one_work_item(ctx):
    """one verdict over 5 blobs x 128 cells (640 cells: more than one 512-cell work item of the bucket MSM and
    of the interpolation), cells of different blobs interleaved so commitment rows are not contiguous"""
    blobs = [oracle_lib.rand_blob((80 + b) << 20) for b in range(5)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs = [], [], [], []
    for i in range(128):
        for b, (st, cl, pr) in enumerate(full):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
    n = len(cells)
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, [0, n]) == [0]
    bad = list(proofs); bad[600] = proofs[601]                       # a valid point, wrong proof, in the second work item
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, [0, n]) == [1]
    # same cells split into two verdicts at an odd boundary; the corrupted proof lands in the second
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, [0, 517, n]) == [0, 1]
    # oracle agrees on the accept case
    o = oracle_lib.get_oracle()
    assert o.verify_cell_kzg_proof_batch(commitments[:130], idx[:130], cells[:130], proofs[:130]) == 0


def test_large_verdict_uses_column_grouping(ctx):
    """a verdict with >= 4096 cells takes the 8-bit-window, grouped-by-column path (vmsm.cuh: k_cell_columns_large); it
    must agree with the small-verdict path on accept and reject, alone and mixed with small verdicts in one call"""
    nblob = 33
    blobs = [oracle_lib.rand_blob((200 + b) << 20) for b in range(nblob)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs = [], [], [], []
    for b, (st, cl, pr) in enumerate(full):
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
    n = len(cells)                                                   # 4224 cells
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, [0, n]) == [0]
    # a large verdict (first 32 blobs = 4096 cells) followed by a small one (last blob) and an empty one
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, [0, 4096, n, n]) == [0, 0, 0]
    bad = list(proofs); bad[1234] = proofs[1235]
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, [0, 4096, n]) == [1, 0]
    badc = list(cells); badc[4100] = cells[4101]
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, badc, proofs, [0, 4096, n]) == [0, 1]
    assert ctx.verify_cell_kzg_proof_batches(commitments, idx, badc, proofs, [0, n]) == [1]
    # wrong cell index on a correct cell/proof pair: the column twiddle must notice
    widx = list(idx); widx[77], widx[78] = widx[78], widx[77]
    assert ctx.verify_cell_kzg_proof_batches(commitments, widx, cells, proofs, [0, n]) == [1]
    # shuffled order inside the verdict (columns gathered from all over the batch) still verifies
    import random
    perm = list(range(n)); random.Random(5).shuffle(perm)
    assert ctx.verify_cell_kzg_proof_batches([commitments[i] for i in perm], [idx[i] for i in perm], [cells[i] for i in perm],
                                             [proofs[i] for i in perm], [0, n]) == [0]


def test_subgroup_check_matches_oracle(ctx):
    """random x-coordinates: on-curve points outside G1 must be rejected exactly as the oracle's [r]P test does"""
    import ctypes, random
    rng = random.Random(7)
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    L = oracle_lib.lib()
    pts = []
    for _ in range(24):
        x = rng.randrange(P)
        b = bytearray(x.to_bytes(48, "big")); b[0] |= 0x80 | (0x20 if rng.random() < 0.5 else 0)
        pts.append(bytes(b))
    pts += [oracle_lib.load_setup()[0][48 * i:48 * i + 48] for i in range(4)]
    exp = []
    for p in pts:
        out = ctypes.create_string_buffer(96)
        exp.append(L.ko_g1_decompress(p, out, 1))
    # drive the GPU decode+subgroup path through VerifyKZGProof: a bad commitment yields its decode status
    zero = bytes(32); inf = bytes([0xc0]) + bytes(47)
    got = ctx.verify_kzg_proof_batch(pts, [zero] * len(pts), [zero] * len(pts), [inf] * len(pts))
    for g, e in zip(got, exp):
        assert (g in (0, 1)) == (e == 0), (g, e)
        if e != 0:
            assert g == e
    assert any(e == 5 for e in exp) and any(e == 4 for e in exp) and any(e == 0 for e in exp)


def test_codec_validation_entry_points(ctx):
    """kzgb200_check_g1_points / kzgb200_check_scalars against the oracle's decoder and Python integers"""
    import ctypes, random
    rng = random.Random(31)
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    m = oracle_lib.load_setup()[0]
    pts = [m[48 * i:48 * i + 48] for i in range(6)] + [bytes([0xc0]) + bytes(47), bytes(48), bytes([0xc0]) + bytes(46) + b"\x01"]
    for _ in range(16):
        b = bytearray(rng.randrange(P).to_bytes(48, "big")); b[0] |= 0x80 | (0x20 if rng.random() < 0.5 else 0)
        pts.append(bytes(b))
    L = oracle_lib.lib()
    exp = [L.ko_g1_decompress(p, ctypes.create_string_buffer(96), 1) for p in pts]
    assert ctx.check_g1_points(pts) == exp and {0, 3, 4, 5} <= set(exp)
    assert ctx.check_g1_points([]) == []
    sc = [0, 1, R - 1, R, R + 1, 2**256 - 1, 2**255] + [rng.randrange(2**256) for _ in range(50)]
    assert ctx.check_scalars([v.to_bytes(32, "big") for v in sc]) == [0 if v < R else 2 for v in sc]
    blob_ok = oracle_lib.rand_blob(9 << 20)
    bad = bytearray(blob_ok); bad[32 * 4095:32 * 4096] = R.to_bytes(32, "big")
    assert ctx.check_scalars([blob_ok, bytes(bad), blob_ok], 4096) == [0, 2, 0]


def test_cell_verifier_reports_first_error_in_reference_order(ctx):
    """api_eip7594.go:184-213: cell indices, then every unique commitment (row order), then every proof (index order), then
    every cell (index order) -- the FIRST failing element's error is the verdict, not the numerically largest status"""
    import ctypes, random
    rng = random.Random(11)
    P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    L = oracle_lib.lib()
    by_status = {}
    while not (4 in by_status and 5 in by_status):
        b = bytearray(rng.randrange(P).to_bytes(48, "big")); b[0] |= 0x80
        st = L.ko_g1_decompress(bytes(b), ctypes.create_string_buffer(96), 1)
        by_status.setdefault(st, bytes(b))
    off_curve, off_subgroup, bad_enc = by_status[4], by_status[5], bytes([0xff]) * 48
    blob = oracle_lib.rand_blob(21 << 20)
    cm = ctx.blob_to_kzg_commitment(blob)[1]
    _, cells, proofs = ctx.compute_cells_and_kzg_proofs(blob)
    cl = [cells[2048 * i:2048 * i + 2048] for i in range(128)]; pl = [proofs[48 * i:48 * i + 48] for i in range(128)]
    idx = list(range(128))
    bad_cell = bytes([0xff]) * 32 + cl[5][32:]

    def run(cms=None, cells_=None, proofs_=None, idx_=None):
        return ctx.verify_cell_kzg_proof_batch(cms or [cm] * 128, idx_ or idx, cells_ or cl, proofs_ or pl)

    assert run() == 0
    # two bad proofs: the earlier one's status wins, whichever is numerically larger
    p = list(pl); p[3] = bad_enc; p[10] = off_subgroup
    assert run(proofs_=p) == 3
    p = list(pl); p[3] = off_subgroup; p[10] = bad_enc
    assert run(proofs_=p) == 5
    # a bad proof late in the batch beats a bad cell early in the batch (all proofs are decoded before any cell)
    c = list(cl); c[5] = bad_cell
    p = list(pl); p[100] = off_curve
    assert run(cells_=c, proofs_=p) == 4
    assert run(cells_=c) == 2
    # a bad commitment beats everything but the cell-index check
    cms = [cm] * 128; cms[127] = off_curve
    p = list(pl); p[0] = off_subgroup
    assert run(cms=cms, cells_=c, proofs_=p) == 4
    cms = [cm] * 128; cms[60] = off_subgroup; cms[61] = bad_enc            # two bad rows: the first-seen row first
    assert run(cms=cms) == 5
    i2 = list(idx); i2[77] = 128
    assert run(cms=cms, cells_=c, proofs_=p, idx_=i2) == 7


def test_optimistic_combined_pass_agrees_with_per_verdict_checks(ctx):
    """many small verdicts (>= 8192 cells in all) are first checked as ONE combined random linear combination
    (kzgb200_verify.cu: optimistic pass); whatever it finds, the per-verdict results must be exactly those of the
    per-verdict checks alone (tunable optimistic = 0): all OK, one false verdict, error statuses in the right slot"""
    nblob = 66
    blobs = [oracle_lib.rand_blob((300 + b) << 20) for b in range(nblob)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs = [], [], [], []
    for b, (st, cl, pr) in enumerate(full):
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
    n = len(cells)                                                   # 8448 cells, 66 verdicts of 128 + one empty
    offs = [128 * b for b in range(nblob + 1)] + [n]
    cases = {"valid": (commitments, idx, cells, proofs)}
    bad = list(proofs); bad[128 * 40 + 7] = proofs[128 * 40 + 8]
    cases["false_proof"] = (commitments, idx, cells, bad)
    badc = list(cells); badc[128 * 65 + 127] = cells[128 * 65 + 126]
    cases["false_cell_last"] = (commitments, idx, badc, proofs)
    bidx = list(idx); bidx[128 * 3 + 1] = 128
    cases["bad_index"] = (commitments, bidx, cells, bad)
    benc = list(proofs); benc[128 * 12] = bytes([0xff]) * 48
    cases["bad_encoding"] = (commitments, idx, cells, benc)
    expect = {"valid": {}, "false_proof": {40: 1}, "false_cell_last": {65: 1}, "bad_index": {3: 7, 40: 1}, "bad_encoding": {12: 3}}
    for name, (cm_, ix_, ce_, pr_) in cases.items():
        got = {}
        for opt in (1, 0):
            assert ctx.L.kzgb200_dbg_set_tunable(b"optimistic", opt) == 0
            try:
                got[opt] = ctx.verify_cell_kzg_proof_batches(cm_, ix_, ce_, pr_, offs)
            finally:
                ctx.L.kzgb200_dbg_set_tunable(b"optimistic", 1)
        want = [expect[name].get(b, 0) for b in range(nblob + 1)]
        assert got[1] == got[0] == want, (name, got)
    # the oracle agrees on the verdict that was made false
    o = oracle_lib.get_oracle()
    lo = 128 * 40
    assert o.verify_cell_kzg_proof_batch(commitments[lo:lo + 128], idx[lo:lo + 128], cells[lo:lo + 128], bad[lo:lo + 128]) == 1
    assert o.verify_cell_kzg_proof_batch(commitments[lo:lo + 128], idx[lo:lo + 128], cells[lo:lo + 128], proofs[lo:lo + 128]) == 0


def test_large_verdict_window_variants_agree(ctx):
    """large verdicts: the 4-bit-window column MSM (default) against the round-1 8-bit-window form (tunable large_window = 8)"""
    nblob = 40
    blobs = [oracle_lib.rand_blob((400 + b) << 20) for b in range(nblob)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs = [], [], [], []
    for b, (st, cl, pr) in enumerate(full):
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
    n = len(cells)
    bad = list(proofs); bad[3000] = proofs[3001]
    for lw in (4, 8):
        assert ctx.L.kzgb200_dbg_set_tunable(b"large_window", lw) == 0
        try:
            assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, proofs, [0, n]) == [0]
            assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, [0, n]) == [1]
            assert ctx.verify_cell_kzg_proof_batches(commitments, idx, cells, bad, [0, 4096, n]) == [1, 0]
        finally:
            ctx.L.kzgb200_dbg_set_tunable(b"large_window", 4)


def test_tail_overlap_split_agrees_with_one_lane(ctx):
    """(an option, off by default: measured without gain) a big share of VerifyCellKZGProofBatch is cut 7/8 + 1/8 onto two lanes of one GPU, the small part gated behind the big part's decode
    (kzgb200_api.cu, LaneGate).  With the threshold lowered so that 10 240 cells are cut (8 960 + 1 280: the big part still takes the optimistic pass), results must equal the one-lane results: many verdicts
    (all valid; a false verdict in the big part; one in the small part; an error) and one large verdict (valid, false in either part)."""
    nblob = 80
    blobs = [oracle_lib.rand_blob((600 + b) << 20) for b in range(nblob)]
    cms = [c for _, c in ctx.blob_to_kzg_commitment_batch(blobs)]
    full = ctx.compute_cells_and_kzg_proofs_batch(blobs)
    commitments, idx, cells, proofs = [], [], [], []
    for b, (st, cl, pr) in enumerate(full):
        for i in range(128):
            commitments.append(cms[b]); idx.append(i); cells.append(cl[2048 * i:2048 * (i + 1)]); proofs.append(pr[48 * i:48 * (i + 1)])
    n = len(cells)
    offs = [128 * b for b in range(nblob + 1)]
    bad_big = list(proofs); bad_big[128 * 10 + 3] = proofs[128 * 10 + 4]
    bad_small = list(cells); bad_small[128 * 78 + 100] = cells[128 * 78 + 99]            # verdict 78 of 80 lies in the last eighth (the cut is at verdict 70)
    bad_enc = list(proofs); bad_enc[128 * 79 + 5] = bytes([0xff]) * 48
    cases = [(proofs, cells, {}), (bad_big, cells, {10: 1}), (proofs, bad_small, {78: 1}), (bad_enc, cells, {79: 3}), (bad_big, bad_small, {10: 1, 78: 1})]
    for tail in (1, 0):
        assert ctx.L.kzgb200_dbg_set_tunable(b"tail_split", tail) == 0 and ctx.L.kzgb200_dbg_set_tunable(b"tail_min_cells", 2048) == 0
        try:
            for pr_, ce_, want in cases:
                assert ctx.verify_cell_kzg_proof_batches(commitments, idx, ce_, pr_, offs) == [want.get(b, 0) for b in range(nblob)], (tail, want)
                one = ctx.verify_cell_kzg_proof_batches(commitments, idx, ce_, pr_, [0, n])
                assert one == [max(want.values()) if want else 0] or (3 in want.values() and one == [3]), (tail, want, one)
        finally:
            ctx.L.kzgb200_dbg_set_tunable(b"tail_split", 0); ctx.L.kzgb200_dbg_set_tunable(b"tail_min_cells", 128 << 10)
