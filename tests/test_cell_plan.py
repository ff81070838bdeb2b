"""Host bookkeeping of VerifyCellKZGProofBatch (csrc/cell_plan.hpp) against a Python model.  No GPU: the hook runs on the host.
Reference behaviour modelled: per-verdict de-duplication of commitments on raw bytes in first-seen order
(api_eip7594.go:238-265), cell indices < 128 (api_eip7594.go:184-188)."""
import ctypes, json, os, random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    L = ctypes.CDLL(os.path.join(ROOT, "go-eth-kzg_b200", "libkzgb200.so"))
    L.kzgb200_dbg_plan_cell_batches_json.restype = ctypes.c_char_p
    L.kzgb200_dbg_plan_cell_batches_json_l.restype = ctypes.c_char_p
    return L


def plan(commitments, idx, offs, item=512, large=4096, row_item=64, l_item=None):
    n, nb = len(commitments), len(offs) - 1
    a = (ctypes.c_uint64 * max(n, 1))(*idx)
    o = (ctypes.c_uint64 * (nb + 1))(*offs)
    u64 = ctypes.c_uint64
    if l_item is None:
        r = _lib().kzgb200_dbg_plan_cell_batches_json(b"".join(commitments), a, ctypes.c_size_t(n), o, ctypes.c_size_t(nb), u64(item), u64(large), u64(row_item))
    else:
        r = _lib().kzgb200_dbg_plan_cell_batches_json_l(b"".join(commitments), a, ctypes.c_size_t(n), o, ctypes.c_size_t(nb), u64(item), u64(large), u64(row_item), u64(l_item))
    return None if r is None else json.loads(r)


def model(commitments, idx, offs, item, large, row_item):
    nb = len(offs) - 1
    P = {k: [] for k in ("bstatus", "batch_start", "large_of", "uniq", "rows", "row_batch", "items", "vs_items", "large_ids", "l_groups", "r_items")}
    P["batch_row_off"], P["batch_item_off"], P["vs_batch_item_off"] = [0], [0], [0]
    batch_of = [0] * len(commitments)
    for b in range(nb):
        lo, hi = offs[b], offs[b + 1]
        P["batch_start"].append(lo)
        seen, rows = {}, []
        bad = 0
        for k in range(lo, hi):
            batch_of[k] = b
            c = commitments[k]
            if c not in seen:
                seen[c] = len(rows); rows.append([]); P["uniq"].append(c)
            rows[seen[c]].append(k)
            if idx[k] >= 128:
                bad = 7
        P["bstatus"].append(bad)
        first_row = len(P["rows"])
        P["rows"] += rows; P["row_batch"] += [b] * len(rows)
        P["batch_row_off"].append(len(P["rows"]))
        runs = [(s, min(hi, s + item)) for s in range(lo, hi, item)]
        P["items"] += runs; P["batch_item_off"].append(len(P["items"]))
        if hi - lo >= large:
            P["large_of"].append(len(P["large_ids"])); P["large_ids"].append(b)
            P["r_items"].append([(s, min(len(P["rows"]), s + row_item)) for s in range(first_row, len(P["rows"]), row_item)])
            P["l_groups"].append([[k for k in range(lo, hi) if (idx[k] & 127) == q] for q in range(128)])
        else:
            P["large_of"].append(-1); P["vs_items"] += runs
        P["vs_batch_item_off"].append(len(P["vs_items"]))
    P["batch_of"] = batch_of
    return P


def check(commitments, idx, offs, item=512, large=4096, row_item=64, l_item=None):
    got, exp = plan(commitments, idx, offs, item, large, row_item, l_item), model(commitments, idx, offs, item, large, row_item)
    l_run = l_item or item                                             # run length of the large verdicts' column items
    assert got is not None
    for k in ("bstatus", "batch_start", "batch_row_off", "batch_item_off", "vs_batch_item_off", "large_of", "large_ids", "batch_of", "row_batch"):
        assert got[k] == exp[k], k
    assert bytes(got["uniq_bytes"]) == b"".join(exp["uniq"])
    rows = [got["row_cells"][got["row_off"][r]:got["row_off"][r + 1]] for r in range(len(got["row_off"]) - 1)]
    assert rows == exp["rows"]                                         # cells of a row in their original order
    assert list(zip(got["item_start"], got["item_end"])) == exp["items"]
    assert list(zip(got["vs_item_start"], got["vs_item_end"])) == exp["vs_items"]
    # large verdicts: column groups (original order inside a column) cut into runs of <= item; slot = large index * 128 + column
    slot = 0
    for lb, groups in enumerate(exp["l_groups"]):
        for q, g in enumerate(groups):
            its = range(got["l_slot_item_off"][slot], got["l_slot_item_off"][slot + 1])
            cells = []
            for it in its:
                s, e = got["l_item_start"][it], got["l_item_end"][it]
                assert 0 < e - s <= l_run
                cells += got["l_order"][s:e]
            assert cells == g, (lb, q)
            slot += 1
        rit = range(got["r_slot_item_off"][lb], got["r_slot_item_off"][lb + 1])
        assert [(got["r_item_start"][i], got["r_item_end"][i]) for i in rit] == exp["r_items"][lb]
    assert slot + 1 == len(got["l_slot_item_off"])
    return got


def test_plan_matches_model():
    rng = random.Random(9)
    cm = [bytes([i]) * 48 for i in range(12)]
    # the layout of the bench: verdicts of 128 cells, one commitment each
    check([cm[b] for b in range(4) for _ in range(128)], [i for _ in range(4) for i in range(128)], [0, 128, 256, 384, 512])
    # interleaved commitments, repeated cells, an empty verdict, a bad index in one verdict only
    n = 700
    c2 = [cm[rng.randrange(5)] for _ in range(n)]
    i2 = [rng.randrange(128) for _ in range(n)]; i2[650] = 128
    got = check(c2, i2, [0, 0, 300, 301, 640, n], item=64)
    assert got["bstatus"] == [0, 0, 0, 0, 7]
    # the same commitment in two verdicts is a row in each of them (de-duplication is per call of the reference, i.e. per verdict)
    got = check([cm[0]] * 6, [1, 2, 3, 1, 2, 3], [0, 3, 6])
    assert len(got["row_off"]) - 1 == 2
    # a large verdict (columns gathered from all over it) next to small ones, small thresholds to exercise the run cutting
    n = 1000
    c3 = [cm[rng.randrange(7)] for _ in range(n)]
    i3 = [rng.randrange(128) for _ in range(n)]
    got = check(c3, i3, [0, 50, 900, 1000], item=16, large=200, row_item=2)
    assert got["large_of"] == [-1, 0, -1]
    # the engine's shape: long interpolation runs, short column runs for the large verdicts (kzgb200_verify.cu: ITEM = 512, L_ITEM = 128)
    got = check(c3, i3, [0, 50, 900, 1000], item=64, large=200, row_item=2, l_item=4)
    assert max(e - s for s, e in zip(got["l_item_start"], got["l_item_end"])) == 4
    # nothing at all, and malformed offsets
    check([], [], [0, 0])
    assert plan([cm[0]] * 4, [0, 1, 2, 3], [0, 5]) is None
    assert plan([cm[0]] * 4, [0, 1, 2, 3], [3, 2]) is None
    # offsets that leave cells outside every verdict are malformed too (such cells used to be decoded anyway and
    # their errors folded into verdict 0)
    assert plan([cm[0]] * 4, [0, 1, 2, 3], [1, 4]) is None
    assert plan([cm[0]] * 4, [0, 1, 2, 3], [0, 3]) is None
    assert plan([cm[0]] * 4, [0, 1, 2, 3], [0, 2, 3]) is None
