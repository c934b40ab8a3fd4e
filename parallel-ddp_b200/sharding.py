"""Sharding over ranks.

* problems: independent, contiguous shards, no data-path collective (shard_range, gather_counters);
* step sizes of one problem's line search: each rank simulates a contiguous range of the alpha candidates, one exchange at selection
  (alpha_range, merge_selection_inputs -- the host-side statement of what libpddp's pddp_alpha_shard_* does on the device)."""
import numpy as np


def shard_range(rank, world, global_batch):
    """Contiguous problem range [lo, hi) of `rank`; sizes differ by at most one."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_counters(local, group=None):
    """all_gather of a 1-D int32 tensor of per-problem counters; returns them in problem order.  Shards may differ in size by one
    (shard_range): every rank pads to the largest shard, the padding is cut after the gather."""
    import torch.distributed as dist
    import torch
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = torch.zeros(max(sizes), dtype=local.dtype, device=local.device); pad[:local.numel()] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)])


def alpha_range(rank, world, n_alpha):
    """Step sizes [a_first, a_first + a_cnt) simulated by `rank` (n_alpha must be a multiple of world)."""
    if n_alpha % world:
        raise ValueError("n_alpha must be a multiple of the number of ranks")
    cnt = n_alpha // world
    return rank * cnt, cnt


def merge_selection_inputs(local_pairs, group=None):
    """all_gather of this rank's (J, defect) pairs [B, a_cnt, 2] -> [B, n_alpha, 2] in step-size order: the one exchange of the sharded
    line search.  Every rank then runs the same sequential scan (fpHelpers.cuh:395-408) on the same data."""
    import torch.distributed as dist
    import torch
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local_pairs) for _ in range(world)]
    dist.all_gather(parts, local_pairs.contiguous(), group=group)
    return torch.cat(parts, dim=1)
