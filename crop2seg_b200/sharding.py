"""Patch sharding for multi-GPU inference.

Every output pixel depends only on its own sample, so the path shards by patch with no data-path
collective: rank r of N owns a contiguous block of the patch list.
"""
from __future__ import annotations

from typing import Sequence, Tuple, TypeVar

T = TypeVar("T")


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of rank's contiguous block; the first ``n_items % world_size`` ranks get one extra."""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_patches(items: Sequence[T], rank: int, world_size: int) -> Sequence[T]:
    """The slice of ``items`` (a tensor batch or a list of patch ids) owned by ``rank``."""
    lo, hi = shard_bounds(len(items), rank, world_size)
    return items[lo:hi]


def gather_shards(local, world_size: int, group=None):
    """Concatenate per-rank result tensors (dim 0) on every rank, in rank order.

    Shards may differ by one item, so they are exchanged with ``all_gather_object``-free padding:
    every rank pads to the largest shard, ``all_gather``s, and trims.  This is the only collective of
    the inference path and it moves results (class maps), never features.
    """
    import torch
    import torch.distributed as dist

    if world_size == 1:
        return local
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world_size)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    n_max = max(sizes)
    padded = local
    if local.shape[0] < n_max:
        pad = torch.zeros((n_max - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
