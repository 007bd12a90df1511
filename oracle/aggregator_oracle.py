"""Numpy restatement of the reference TemporalAggregator.  TEST INFRASTRUCTURE ONLY.

    TemporalAggregator.forward        src/backbones/temporal_aggregator.py:14-77
    nn.Upsample(bilinear, align_corners=False)   called at temporal_aggregator.py:17-19,27
    nn.AvgPool2d(kernel_size=w // H)             called at temporal_aggregator.py:29

The bilinear rule is ATen's ``area_pixel_compute_source_index``: with
``scale = in / out`` (float32), ``src = scale * (dst + 0.5) - 0.5`` clamped at 0,
``i0 = floor(src)``, ``i1 = min(i0 + 1, in - 1)``, ``l1 = src - i0``, ``l0 = 1 - l1``.

Parity status: pinned against outputs of the imported reference
(``tests/golden/make_golden.py``, ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

F32 = np.float32


def _source_index(out_size: int, in_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    scale = F32(in_size) / F32(out_size)
    dst = np.arange(out_size, dtype=F32)
    src = (scale * (dst + F32(0.5)) - F32(0.5)).astype(F32)
    src = np.maximum(src, F32(0.0))
    i0 = np.floor(src).astype(np.int64)
    i0 = np.minimum(i0, in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(F32)).astype(F32)
    l0 = (F32(1.0) - l1).astype(F32)
    return i0, i1, l0, l1


def bilinear_upsample(a: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """Bilinear resize of the last two axes of ``a`` to ``size`` (align_corners=False)."""
    a = a.astype(F32)
    hi, wi = a.shape[-2:]
    ho, wo = size
    y0, y1, ly0, ly1 = _source_index(ho, hi)
    x0, x1, lx0, lx1 = _source_index(wo, wi)
    top = a[..., y0, :]
    bot = a[..., y1, :]
    t = top[..., :, x0] * lx0 + top[..., :, x1] * lx1
    b = bot[..., :, x0] * lx0 + bot[..., :, x1] * lx1
    return (t * ly0[:, None] + b * ly1[:, None]).astype(F32)


def avg_pool2d(a: np.ndarray, k: int) -> np.ndarray:
    """``nn.AvgPool2d(kernel_size=k)`` (stride k, no padding, floor) on the last two axes."""
    hi, wi = a.shape[-2:]
    ho, wo = hi // k, wi // k
    a = a[..., : ho * k, : wo * k].astype(F32)
    a = a.reshape(a.shape[:-2] + (ho, k, wo, k))
    return a.mean(axis=(-3, -1), dtype=np.float64).astype(F32)


def temporal_aggregator(x: np.ndarray, pad_mask: Optional[np.ndarray] = None,
                        attn_mask: Optional[np.ndarray] = None, mode: str = "mean") -> np.ndarray:
    """``TemporalAggregator(mode).forward(x, pad_mask, attn_mask)`` (temporal_aggregator.py:14-77).

    x[B, T, C, H, W]; attn_mask[h, B, T, ha, wa]; pad_mask[B, T] bool.  The masked branch
    (taken when any frame is padded) multiplies the resized attention by ``~pad_mask``.
    """
    x = x.astype(F32)
    b, t, c, hh, ww = x.shape
    masked = pad_mask is not None and bool(np.any(pad_mask))
    keep = None
    if masked:
        keep = (~pad_mask.astype(bool)).astype(F32)  # [B, T]

    if mode == "att_group":
        n_heads, _, _, ha, wa = attn_mask.shape
        if hh > wa:  # temporal_aggregator.py:26-29 compares x's height with the attention width
            attn = bilinear_upsample(attn_mask, (hh, ww))
        else:
            attn = avg_pool2d(attn_mask, wa // hh)
        if masked:
            attn = attn * keep[None, :, :, None, None]  # :33
        xg = x.reshape(b, t, n_heads, c // n_heads, hh, ww)  # :35 chunk into head groups
        out = np.einsum("hbtyx,bthcyx->bhcyx", attn, xg, dtype=F32)  # :37-38
        return out.reshape(b, c, hh, ww).astype(F32)  # :44
    if mode == "att_mean":
        attn = attn_mask.astype(F32).mean(axis=0, dtype=np.float64).astype(F32)  # :48 / :72
        attn = bilinear_upsample(attn, (hh, ww))
        if masked:
            attn = attn * keep[:, :, None, None]
        return np.einsum("btyx,btcyx->bcyx", attn, x, dtype=F32).astype(F32)
    if mode == "mean":
        if masked:  # :53-56
            s = (x * keep[:, :, None, None, None]).sum(axis=1, dtype=np.float64)
            return (s / keep.sum(axis=1, dtype=np.float64)[:, None, None, None]).astype(F32)
        return x.mean(axis=1, dtype=np.float64).astype(F32)  # :77
    raise ValueError(f"unknown aggregation mode {mode!r}")


def pad_mask_from_input(x: np.ndarray, pad_value: float = 0.0) -> np.ndarray:
    """``(input == pad_value).all(dim=-1).all(dim=-1).all(dim=-1)`` -- utae.py:201-203, wtae.py:221-223,
    timeunet.py:170-172: bool [B,T], True where every element of the frame equals ``pad_value`` (NaN never does)."""
    return (x == x.dtype.type(pad_value)).all(axis=(-1, -2, -3))
