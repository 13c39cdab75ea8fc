"""Batch sharding across the GPUs of one box: inference and training both split the
batch into contiguous per-rank shards; inference needs no collective (results are concatenated by the
caller), training sums the flat gradient buffer with ONE all-reduce per step."""
from __future__ import annotations

from typing import Tuple

import torch


def batch_shard(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `total` items for `rank`; sizes differ by at most one, low ranks get the
    extra items, and with total < world the high ranks get empty shards (fewer GPUs are used)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} / world {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def all_reduce_sum(flat: torch.Tensor) -> torch.Tensor:
    """The training step's single exchange: in-place SUM of the flat fp32 gradient buffer over all ranks
    (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests).  The local loss is pre-scaled by 1/world, so the
    sum is the gradient of the global-batch mean (per-replica BatchNorm statistics, like the reference's
    DataParallel)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat
