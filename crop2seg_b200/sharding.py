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


class GradientBucket:
    """One flat fp32 buffer for every gradient of ``params``, so that the data-parallel training step of the hot path
    needs exactly ONE collective: ``all_reduce()`` packs the gradients autograd produced into the buffer with one
    multi-tensor copy, averages the buffer over the ranks in place (NCCL over NVLink / NVSwitch) and leaves every
    ``p.grad`` as a view into it -- no per-parameter hooks, no bucket bookkeeping on the host.  SURVEY.md section 8e:
    training = replicas + one gradient all-reduce; BatchNorm statistics stay per replica like in the reference (train.py
    builds no SyncBatchNorm).

        bucket = GradientBucket(encoder.parameters())
        bucket.zero(); loss.backward(); bucket.all_reduce(); optimizer.step()

    ``zero()`` drops the gradients (``p.grad = None``): autograd then hands its tensors over instead of ADDING them into
    existing ones -- with gradients pre-attached as views that was one small ``add_`` launch per parameter and step (16 for
    the encoder) plus a memset.  On a single rank ``all_reduce()`` is a no-op and nothing is copied at all.
    """

    def __init__(self, params, group=None):
        import torch
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradientBucket: no trainable parameter")
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.group = group
        self.views = []
        off = 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("GradientBucket: parameters must be float32 on one device")
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def zero(self) -> None:
        for p in self.params:
            p.grad = None

    def pack(self) -> None:
        """Gather the parameters' gradients into the flat buffer (one multi-tensor copy) and re-point them at it."""
        import torch
        src, dst, missing = [], [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                missing.append(v)
            elif g.data_ptr() != v.data_ptr():
                src.append(g.detach().to(torch.float32)), dst.append(v)
        if missing:
            torch._foreach_zero_(missing)
        if dst:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self) -> None:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        self.pack()
        dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM,
                        group=self.group)
        if dist.get_backend(self.group) != "nccl":  # gloo has no AVG
            self.flat.div_(dist.get_world_size(self.group))
