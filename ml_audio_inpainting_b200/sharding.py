"""Clip sharding across the GPUs of one box: one process per GPU, contiguous blocks of the batch,
NO collective on the data path (clips are independent).

Random draws (gap starts) are made ONCE for the whole batch from the global ``np.random`` stream, in the
reference's order, on every rank identically (same seed => same stream), and each rank slices its block --
so results do not depend on the number of GPUs.  ``gather_rows`` is the optional final gather
(torch.distributed all_gather: NCCL over NVLink on GPUs, gloo on CPU tensors in the tests).
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

__all__ = ["world", "shard_bounds", "shard_slice", "gather_rows"]


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank ``rank``: the first ``n % world`` ranks get one extra item."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_slice(array, rank: int, world_size: int):
    lo, hi = shard_bounds(len(array), rank, world_size)
    return array[lo:hi]


def gather_rows(local, n_items: int, group=None):
    """All-gather the per-rank row blocks (possibly ragged by one row) back into [n_items, ...] on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    ws = dist.get_world_size(group)
    rows = max(shard_bounds(n_items, r, ws)[1] - shard_bounds(n_items, r, ws)[0] for r in range(ws))
    pad = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    parts = []
    for r in range(ws):
        lo, hi = shard_bounds(n_items, r, ws)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, 0)
