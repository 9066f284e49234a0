"""Multi-GPU harness for the inference path: the path shards over independent images, so N
GPUs run N replicas on disjoint shards with NO collective on the data path; only the timing is reduced
(max over ranks).  Training's one collective is DDP's gradient all-reduce (utils/torch_utils.smart_DDP)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` units for `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def aggregate_throughput(images: int, ms: float, device) -> tuple[int, float]:
    """(sum of units over ranks, max of elapsed ms over ranks); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return images, ms
    t = torch.tensor([float(images)], device=device, dtype=torch.float64)
    w = torch.tensor([float(ms)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    return int(round(t.item())), float(w.item())
