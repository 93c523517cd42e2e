"""Stream sharding across the GPUs of one box (one process per GPU, ``torch.distributed``).

Streams share nothing but read-only weights, so the data path needs **no collective**: every rank
runs its own contiguous range of stream ids on its own engine.  The helpers below are for batch
front-ends where one rank owns the I/O: contiguous partition, the one-off broadcast of the packed weight
blob at start-up (every rank holds a full copy, <= 17 MB), the scatter of a hop's PCM from the fronting
rank and the gather of the enhanced PCM (NCCL over NVLink on GPUs, gloo in the CPU tests).
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


def broadcast_weights(blob, src: int = 0, device=None, group=None) -> bytes:
    """Start-up: the packed weight blob (``weights.pack_checkpoint``) of rank ``src`` on every rank.

    Only ``src`` needs the checkpoint file; the others pass ``None``.  One size broadcast, one payload broadcast."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = device if device is not None else "cpu"
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src, group=group)
    if rank == src:
        buf = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    else:
        buf = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def scatter_rows(full, n_streams: int, src: int = 0, group=None):
    """The fronting rank ``src`` holds ``[n_streams, ...]``; every rank receives its own contiguous row block
    (grouped point-to-point sends: 640 B per stream and hop, no collective on the data path otherwise)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts = partition(n_streams, world)
    a, b = parts[rank]
    if rank == src:
        reqs = [dist.isend(full[pa:pb].contiguous(), r, group=group) for r, (pa, pb) in enumerate(parts) if r != src and pb > pa]
        for q in reqs:
            q.wait()
        return full[a:b]
    out = torch.empty((b - a,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)   # `full` on other ranks: a template with the row shape
    if b > a:
        dist.recv(out, src, group=group)
    return out
