"""Numpy restatement of the aggregation followed by the decoder's skip convolution.  TEST INFRASTRUCTURE ONLY.

    skip = TemporalAggregator('att_group')(x, pad_mask, attn)         utae.py:225-227, temporal_aggregator.py:14-45
    UpConvBlock.skip_conv = Conv2d(d, d, 1) -> BatchNorm2d(d) -> ReLU  conv.py:378-382, applied at conv.py:408

Eval mode (running statistics), as in inference.  Parity status: pinned against outputs of the imported reference
(``tests/golden/make_skipconv_golden.py`` -> ``tests/golden/skipconv_*.npz``, checked in tests/test_skipconv.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .aggregator_oracle import temporal_aggregator


def skip_conv(skip: np.ndarray, p: Dict[str, np.ndarray], eps: float = 1e-5) -> np.ndarray:
    """``p`` holds the module's state_dict: 0.weight [d,d,1,1], 0.bias, 1.weight, 1.bias, 1.running_mean, 1.running_var."""
    w = p["0.weight"].reshape(p["0.weight"].shape[0], -1).astype(np.float64)
    y = np.einsum("oc,bchw->bohw", w, skip.astype(np.float64))
    if p.get("0.bias") is not None:
        y = y + p["0.bias"].astype(np.float64)[None, :, None, None]
    y = (y - p["1.running_mean"].astype(np.float64)[None, :, None, None]) / np.sqrt(
        p["1.running_var"].astype(np.float64)[None, :, None, None] + eps)
    y = y * p["1.weight"].astype(np.float64)[None, :, None, None] + p["1.bias"].astype(np.float64)[None, :, None, None]
    return np.maximum(y, 0.0).astype(np.float32)


def aggregate_skip_conv(x: np.ndarray, pad_mask: Optional[np.ndarray], attn: np.ndarray, p: Dict[str, np.ndarray],
                        eps: float = 1e-5, round_skip=None) -> np.ndarray:
    """``round_skip``: optional callable applied to the aggregated skip map (the bf16 path stores it as bf16)."""
    skip = temporal_aggregator(x, pad_mask, attn, "att_group")
    if round_skip is not None:
        skip = round_skip(skip)
    return skip_conv(skip, p, eps)
