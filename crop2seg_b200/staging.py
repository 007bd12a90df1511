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


def smart_forward(forward, x: torch.Tensor, pad_value=None, pad_mask: torch.Tensor = None) -> torch.Tensor:
    """``TemporallySharedBlock.smart_forward`` (temp_shared_block.py:18-47) without its waste: apply ``forward`` (a block
    shared across the sequence, [N, C, H, W] -> [N, C', H', W']) to the non-padded frames of x[B, T, C, H, W] and
    return [B, T, C', H', W'] with ``pad_value`` on the padded frames.

    Differences from the reference, results being identical: the pad mask comes from the early-exit scan kernel
    (``pad_mask_from_input``) or from the caller (``pad_mask`` [B, T], e.g. the one the model already derived from the
    raw input: padded frames stay exactly ``pad_value`` through the encoder, temp_shared_block.py:30-40) instead of
    comparing the whole tensor again; the output shape is taken from the real forward instead of an extra forward of
    an all-zero dummy batch; the valid frames are gathered once by index."""
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
    valid = (~pad_mask.reshape(-1)).nonzero(as_tuple=True)[0]  # the one host synchronisation (the reference has two)
    if valid.numel() == b * t:
        out = forward(flat)
        return out.view(b, t, *out.shape[1:])
    if valid.numel() == 0:
        raise RuntimeError("smart_forward: every frame is padded (the reference fails on the empty batch too)")
    part = forward(flat.index_select(0, valid))
    out = torch.full((b * t, *part.shape[1:]), float(pad_value), dtype=part.dtype, device=part.device)
    out.index_copy_(0, valid, part)
    return out.view(b, t, *out.shape[1:])
