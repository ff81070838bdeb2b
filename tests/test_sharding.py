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
