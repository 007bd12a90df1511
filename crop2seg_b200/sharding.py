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
