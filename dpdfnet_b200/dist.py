"""Stream sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Streams share nothing but read-only weights, so the data path needs **no collective**: every rank
runs its own contiguous range of stream ids on its own engine.  The helpers below are for batch
front-ends where one rank owns the I/O: contiguous partition, and the trivial gather of the enhanced
PCM (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def partition(n_streams: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous ``[start, stop)`` stream-id ranges, sizes differing by at most one."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    base, extra = divmod(int(n_streams), world_size)
    out, start = [], 0
    for r in range(world_size):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def local_range(n_streams: int, rank: int, world_size: int) -> Tuple[int, int]:
    return partition(n_streams, world_size)[rank]


def shard(batch: np.ndarray, rank: int, world_size: int) -> np.ndarray:
    a, b = local_range(batch.shape[0], rank, world_size)
    return batch[a:b]


def gather_rows(local, n_streams: int, group=None):
    """All-gather per-rank row blocks ``[n_local, ...]`` (torch tensor) into ``[n_streams, ...]``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = partition(n_streams, world)
    width = max(b - a for a, b in parts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:b - a] for r, (a, b) in enumerate(parts)], 0)
