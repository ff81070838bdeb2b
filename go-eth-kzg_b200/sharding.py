"""Blob sharding across the GPUs of one box: contiguous blob ranges per rank,
no data-path collective; results come back by a host gather into disjoint slices."""


def shard_range(n_items, world, rank):
    """contiguous [lo, hi) of rank `rank`; sizes differ by at most one, earlier ranks get the extra"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def host_gather(local_bytes, dist, dst=0):
    """gathers variable-length byte strings to rank dst in rank order (torch.distributed, any backend)"""
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local_bytes, out, dst=dst)
    return b"".join(out) if out is not None else None
