"""Problem sharding over ranks: independent problems, contiguous shards, no data-path collective."""
import numpy as np


def shard_range(rank, world, global_batch):
    """Contiguous problem range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_counters(local, group=None):
    """all_gather of a 1-D int32 tensor of per-problem counters (equal shard sizes); returns them in problem order."""
    import torch.distributed as dist
    import torch
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, local, group=group)
    return torch.cat(parts)
