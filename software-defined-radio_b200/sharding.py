"""Capture sharding across ranks (one process per GPU).

Captures are independent -- nothing in src/filter.cpp or src/project.cpp couples two captures --
so the multi-GPU path partitions them by contiguous ranges and needs no collective on the data
path.  The only exchange is the final host-side gather of PCM."""
from __future__ import annotations

import numpy as np


def shard_range(n_captures: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous [begin, end) of captures owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n_captures, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_pcm(local_pcm: np.ndarray, n_captures: int, dist=None, dst: int = 0):
    """Final host gather: rank `dst` receives [n_captures, n_pcm] int16 in capture order, other
    ranks receive None.  `dist` is an initialised torch.distributed module (gloo or nccl group
    with CPU tensors via gloo); with dist=None (single process) the input is returned."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_pcm
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    n_pcm = local_pcm.shape[1]
    sizes = [shard_range(n_captures, world, r) for r in range(world)]
    biggest = max(e - b for b, e in sizes)
    # transported as bytes: every backend moves uint8, not every backend moves int16
    buf = torch.zeros((biggest, n_pcm), dtype=torch.int16)
    buf[: local_pcm.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_pcm))
    raw = buf.view(torch.uint8)
    out = [torch.zeros_like(raw) for _ in range(world)] if rank == dst else None
    dist.gather(raw, out, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([out[r].view(torch.int16)[: e - b].numpy() for r, (b, e) in enumerate(sizes)], axis=0)


def gather_rds(local_reads: list, n_captures: int, dist=None, dst: int = 0):
    """Final host gather of the RDS bit layer: `local_reads` is this rank's list of per-capture
    results (what ``Rds.read(c)`` returns: a dict of small arrays and the offsets string), in
    shard order.  Rank `dst` receives the list for all captures in capture order, other ranks
    None.  The payload is a few hundred bits per capture and block, so it travels as objects."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_reads
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == dst else None
    dist.gather_object(local_reads, out, dst=dst)
    if rank != dst:
        return None
    merged = [r for part in out for r in part]
    if len(merged) != n_captures:
        raise ValueError("shards do not add up to the batch")
    return merged
