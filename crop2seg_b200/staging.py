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
