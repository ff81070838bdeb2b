"""world_size-2 gloo test of the multi-GPU host logic: contiguous blob shards, host gather in rank
order.  The per-blob work is stood in for by the oracle (CPU) -- the point is the plumbing."""
import os, sys
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, os.path.join(ROOT, "go-eth-kzg_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import sharding, oracle_lib
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(n, world, rank)
    o = oracle_lib.get_oracle()
    mine = b"".join(o.blob_to_kzg_commitment(oracle_lib.rand_blob(b << 20))[1] for b in range(lo, hi))
    allc = sharding.host_gather(mine, dist)
    if rank == 0:
        q.put(allc)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition():
    import sharding
    for n in (0, 1, 5, 1024, 1025):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


def test_two_rank_gloo_gather_matches_single_process():
    import oracle_lib
    n = 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=300)
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    o = oracle_lib.get_oracle()
    exp = b"".join(o.blob_to_kzg_commitment(oracle_lib.rand_blob(b << 20))[1] for b in range(n))
    assert got == exp


# ---- the library's own shard planner (csrc/shard_plan.hpp through a host-only debug hook; no GPU) -------------------
def _plan(n_units, n_dev, min_per_dev, offs=None, min_cells=0, sub=None):
    import ctypes, json
    import kzgb200
    L = kzgb200.load_library()
    L.kzgb200_dbg_shard_plan_json.restype = ctypes.c_char_p
    o = (ctypes.c_uint64 * len(offs))(*offs) if offs is not None else None
    s = (ctypes.c_int32 * len(sub))(*sub) if sub else None
    r = L.kzgb200_dbg_shard_plan_json(ctypes.c_size_t(n_units), ctypes.c_size_t(n_dev), ctypes.c_size_t(min_per_dev), o,
                                      ctypes.c_size_t(len(offs) - 1 if offs is not None else 0), ctypes.c_size_t(min_cells), s, ctypes.c_size_t(len(sub) if sub else 0))
    return json.loads(r)


def test_library_unit_ranges_are_a_contiguous_partition():
    for n in (0, 1, 15, 16, 31, 32, 33, 127, 128, 1024, 1025, 4096):
        for d in (1, 2, 3, 4, 8):
            for m in (1, 16, 64):
                r = _plan(n, d, m)["units"]
                if n == 0:
                    assert r == []
                    continue
                assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
                assert len(r) <= d and all(hi > lo for lo, hi in r)
                if len(r) > 1:
                    assert min(hi - lo for lo, hi in r) >= m and max(hi - lo for lo, hi in r) - min(hi - lo for lo, hi in r) <= 1
                else:
                    assert n < 2 * m or d == 1
                # matches the rank-sharding helper the torchrun harness uses when the sizes agree
                import sharding
                if len(r) == d and n % d == 0:
                    assert r == [list(sharding.shard_range(n, d, k)) for k in range(d)]


def test_library_verdict_ranges_balance_cells():
    import random
    rng = random.Random(3)
    for trial in range(200):
        nb = rng.randrange(1, 40)
        sizes = [rng.choice([0, 1, 64, 128, 128, 128, 5000]) for _ in range(nb)]
        offs = [0]
        for s in sizes:
            offs.append(offs[-1] + s)
        d = rng.choice([1, 2, 4, 8])
        r = _plan(0, d, 1, offs, min_cells=rng.choice([1, 128, 4096]))["verdicts"]
        assert r[0][0] == 0 and r[-1][1] == nb and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert 1 <= len(r) <= min(d, nb) and all(hi > lo for lo, hi in r)
    # 4096 equal verdicts over 8 devices: 512 each
    offs = [128 * i for i in range(4097)]
    assert _plan(0, 8, 1, offs, min_cells=4096)["verdicts"] == [[512 * i, 512 * (i + 1)] for i in range(8)]
    # too few cells to split
    assert _plan(0, 8, 1, [0, 128, 256], min_cells=4096)["verdicts"] == [[0, 2]]


def test_library_sub_verdict_merge_is_first_error_else_and():
    assert _plan(0, 1, 1, sub=[0, 0, 0])["merged"] == 0
    assert _plan(0, 1, 1, sub=[0, 1, 0])["merged"] == 1
    assert _plan(0, 1, 1, sub=[1, 0, 3, 5])["merged"] == 3          # an error beats a failed pairing, first error in index order
    assert _plan(0, 1, 1, sub=[0, 5, 3])["merged"] == 5
