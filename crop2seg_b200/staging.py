"""Lengths-aware host -> device staging of padded time series.

The reference moves whole padded batches to the device and then rescans them for ``pad_value``
(``utae.py:201-203``, ``temp_shared_block.py:18-47``).  The kernels of this package never read a padded frame
(the aggregator skips them; the L-TAE does when ``assume_zero_padded`` is set), so the frames behind the end of a
series need not cross PCIe at all: only the ``L_b`` valid frames of every sample are copied.
"""
from __future__ import annotations

from typing import Sequence

import torch


def valid_lengths(pad_mask: torch.Tensor) -> list:
    """Number of leading valid frames per sample of a [B, T] pad mask (True = padded).

    Raises if a valid frame follows a padded one: prefix copies would drop it (use a plain ``copy_`` then)."""
    m = pad_mask.to("cpu", torch.bool)
    lengths = (~m).sum(dim=1)
    t = torch.arange(m.shape[1]).unsqueeze(0)
    if not torch.equal(m, t >= lengths.unsqueeze(1)):
        raise ValueError("pad_mask is not a suffix mask (valid frames after a padded one)")
    return [int(v) for v in lengths]


def copy_valid_frames_(dst: torch.Tensor, src: torch.Tensor, lengths: Sequence[int], zero_rest: bool = False) -> int:
    """Copy ``src[b, :lengths[b]]`` into ``dst[b, :lengths[b]]`` for every sample (``non_blocking``: ``src`` should be
    pinned), on the current stream.  ``dst`` / ``src`` are [B, T, ...] with identical shapes.  The remaining frames of
    ``dst`` keep whatever they held unless ``zero_rest``.  Returns the number of bytes copied."""
    if dst.shape != src.shape or dst.dtype != src.dtype:
        raise ValueError(f"shape/dtype mismatch: {tuple(dst.shape)} {dst.dtype} vs {tuple(src.shape)} {src.dtype}")
    if len(lengths) != dst.shape[0]:
        raise ValueError("one length per sample expected")
    frame_bytes = src[0, 0].numel() * src.element_size()
    copied = 0
    b = 0
    n = dst.shape[0]
    while b < n:  # runs of full-length samples go out as one copy
        L = int(lengths[b])
        if L < 0 or L > dst.shape[1]:
            raise ValueError(f"length {L} outside [0, {dst.shape[1]}]")
        if L == dst.shape[1]:
            e = b
            while e < n and int(lengths[e]) == dst.shape[1]:
                e += 1
            dst[b:e].copy_(src[b:e], non_blocking=True)
            copied += (e - b) * L * frame_bytes
            b = e
            continue
        if L > 0:
            dst[b, :L].copy_(src[b, :L], non_blocking=True)
            copied += L * frame_bytes
        if zero_rest:
            dst[b, L:].zero_()
        b += 1
    return copied


def frame_slots(pad_mask: torch.Tensor):
    """(slot, n_valid) for a pad mask of any shape (True = padded): ``slot`` int32, position of every frame among the
    valid ones (-1 for padded frames), ``n_valid`` a DEVICE int32 scalar -- ``c2s_frame_index``, no host synchronisation."""
    from . import _lib
    from .ops import _require_cuda, _stream
    _require_cuda(pad_mask, "pad_mask")
    m = pad_mask.reshape(-1)
    if m.dtype != torch.bool:
        m = m != 0
    m = m.contiguous().view(torch.uint8)
    slot = torch.empty(m.numel(), dtype=torch.int32, device=m.device)
    count = torch.empty((), dtype=torch.int32, device=m.device)
    with torch.cuda.device(m.device):
        status = _lib.load().c2s_frame_index(m.data_ptr(), m.numel(), slot.data_ptr(), count.data_ptr(), _stream(m.device))
    _lib.check(status, "c2s_frame_index")
    return slot, count


def gather_frames(flat: torch.Tensor, slot: torch.Tensor, n_valid: int) -> torch.Tensor:
    """``flat[~pad_mask]`` for flat[N, ...]: the valid frames packed in order (``c2s_frames_gather``)."""
    from . import _lib
    from .ops import _dtype_code, _require_cuda, _stream
    _require_cuda(flat, "x")
    flat = flat.contiguous()
    packed = torch.empty((n_valid,) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
    with torch.cuda.device(flat.device):
        status = _lib.load().c2s_frames_gather(flat.data_ptr(), slot.data_ptr(), packed.data_ptr(), flat.shape[0],
                                               flat[0].numel(), _dtype_code(flat, "x"), _stream(flat.device))
    _lib.check(status, "c2s_frames_gather")
    return packed


def scatter_frames(packed: torch.Tensor, slot: torch.Tensor, pad_value: float) -> torch.Tensor:
    """``temp = full(pad_value); temp[~pad_mask] = packed`` in one pass (``c2s_frames_scatter``)."""
    from . import _lib
    from .ops import _dtype_code, _require_cuda, _stream
    _require_cuda(packed, "packed")
    packed = packed.contiguous()
    out = torch.empty((slot.numel(),) + tuple(packed.shape[1:]), dtype=packed.dtype, device=packed.device)
    with torch.cuda.device(packed.device):
        status = _lib.load().c2s_frames_scatter(packed.data_ptr(), slot.data_ptr(), out.data_ptr(), slot.numel(),
                                                packed[0].numel(), _dtype_code(packed, "packed"), float(pad_value),
                                                _stream(packed.device))
    _lib.check(status, "c2s_frames_scatter")
    return out


class _Gather(torch.autograd.Function):
    """gather_frames with its adjoint (a scatter with zeros on the padded frames)."""

    @staticmethod
    def forward(ctx, flat, slot, n_valid):
        ctx.slot = slot
        return gather_frames(flat, slot, n_valid)

    @staticmethod
    def backward(ctx, grad):
        return scatter_frames(grad, ctx.slot, 0.0), None, None


class _Scatter(torch.autograd.Function):
    """scatter_frames with its adjoint (a gather: the padded frames are constants)."""

    @staticmethod
    def forward(ctx, packed, slot, pad_value):
        ctx.slot, ctx.n_valid = slot, packed.shape[0]
        return scatter_frames(packed, slot, pad_value)

    @staticmethod
    def backward(ctx, grad):
        return gather_frames(grad, ctx.slot, ctx.n_valid), None, None


def smart_forward(forward, x: torch.Tensor, pad_value=None, pad_mask: torch.Tensor = None, lengths=None) -> torch.Tensor:
    """``TemporallySharedBlock.smart_forward`` (temp_shared_block.py:18-47) without its waste: apply ``forward`` (a block
    shared across the sequence, [N, C, H, W] -> [N, C', H', W']) to the non-padded frames of x[B, T, C, H, W] and
    return [B, T, C', H', W'] with ``pad_value`` on the padded frames.

    Differences from the reference, results being identical: the pad mask comes from the early-exit scan kernel
    (``pad_mask_from_input``) or from the caller (``pad_mask`` [B, T], e.g. the one the model already derived from the
    raw input: padded frames stay exactly ``pad_value`` through the encoder, temp_shared_block.py:30-40) instead of
    comparing the whole tensor again; the output shape is taken from the real forward instead of an extra forward of
    an all-zero dummy batch; the valid frames are packed and put back by two copy kernels driven by a device-side
    scan of the mask (``c2s_frame_index`` / ``c2s_frames_gather`` / ``c2s_frames_scatter``) instead of boolean
    indexing.  ``lengths`` (valid frames per sample, known on the host since ``pad_collate``, src/utils.py:20-66)
    makes the call free of host synchronisations; without it the number of valid frames is read back once, because
    the batch size of ``forward`` has to be known on the host."""
    if x.dim() == 4:
        return forward(x)
    b, t, c, h, w = x.shape
    flat = x.contiguous().view(b * t, c, h, w)
    if pad_value is None:
        out = forward(flat)
        return out.view(b, t, *out.shape[1:])
    if pad_mask is None:
        from .ops import pad_mask_from_input
        pad_mask = pad_mask_from_input(x, pad_value)
    slot, count = frame_slots(pad_mask)
    n_valid = int(sum(int(v) for v in lengths)) if lengths is not None else int(count.item())
    if n_valid == b * t:
        out = forward(flat)
        return out.view(b, t, *out.shape[1:])
    if n_valid == 0:
        raise RuntimeError("smart_forward: every frame is padded (the reference fails on the empty batch too)")
    part = forward(_Gather.apply(flat, slot, n_valid))
    out = _Scatter.apply(part, slot, float(pad_value))
    return out.view(b, t, *out.shape[1:])
