"""Numpy restatement of the edges of the reference's tile inference pipeline.  TEST INFRASTRUCTURE ONLY.

What happens to a Sentinel-2 tile before and after the model in ``src/webapp``:

    patchify          src/helpers/dataset_creator.py:385-388   zero-pad the RAW tile to a multiple of the patch size
                                                               (np.pad constant 0), cut into 128 x 128 patches, row-major
    channel order     src/datasets/s2_ts_cz_crop.py:248, 374   [2, 1, 0, 4, 5, 6, 3, 7, 8, 9] (channels_like_pastis)
    normalisation     src/datasets/s2_ts_cz_crop.py:393-398    (d - mean[c]) / std[c], fp32, AFTER the reordering
                                                               (zero-padded pixels become -mean/std, not 0)
    temporal padding  src/utils.py:14-33 (pad_collate)         frames T .. max_size-1 = pad_value (0), after normalising
    class map         src/webapp/prediction.py:316-333         softmax over classes, first maximum, 128^2 patches put
                                                               back row-major, cropped to the tile size

Parity status: pinned against outputs of the reference's own code (tests/golden/make_tile_golden.py executes
``DatasetCreator._patchify``, ``S2TSCZCropDataset.__getitem__``, ``pad_collate`` and the statements of
``generate_prediction``; tests/test_tile_edges.py compares bit for bit).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

F32 = np.float32
CHANNELS_LIKE_PASTIS = (2, 1, 0, 4, 5, 6, 3, 7, 8, 9)  # s2_ts_cz_crop.py:248


def patch_grid(h: int, w: int, patch: int = 128) -> Tuple[int, int]:
    """Patches per column / row after zero-padding to a multiple of ``patch`` (dataset_creator.py:385-386: the webapp
    pads 1098 -> 1280; a full 10980 tile pads to 11008)."""
    return -(-h // patch), -(-w // patch)


def patchify_normalise(tile: np.ndarray, channels_order, mean: np.ndarray, std: np.ndarray, t_pad: Optional[int] = None,
                       pad_value: float = 0.0, patch: int = 128, patch_begin: int = 0,
                       patch_count: Optional[int] = None, grid: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """tile[T, C, H, W] raw (any real dtype) -> patches[P, T_pad, C, patch, patch] float32.

    patches[p, t, c] = (pad0(tile)[t, order[c], tile rows/cols of patch p] - mean[c]) / std[c] for t < T, pad_value
    behind.  ``patch_begin`` / ``patch_count`` select a contiguous range of the row-major patch list.  ``grid`` = patch
    rows / columns of the padded tile; default: just enough to cover it.  The webapp always pads by 182 pixels
    (dataset_creator.py:386), i.e. grid (10, 10) for its 1098 x 1098 tiles: the last row and column are pure padding."""
    t, c, h, w = tile.shape
    gh, gw = patch_grid(h, w, patch) if grid is None else grid
    t_pad = t if t_pad is None else t_pad
    n = gh * gw
    count = n - patch_begin if patch_count is None else patch_count
    out = np.full((count, t_pad, c, patch, patch), F32(pad_value), dtype=F32)
    order = list(channels_order)
    m = np.asarray(mean, dtype=F32)[None, :, None, None]
    s = np.asarray(std, dtype=F32)[None, :, None, None]
    for i in range(count):
        ph, pw = divmod(patch_begin + i, gw)
        y0, x0 = ph * patch, pw * patch
        raw = np.zeros((t, c, patch, patch), dtype=F32)  # np.pad(..., 'constant') of the RAW values
        ys, xs = max(0, min(patch, h - y0)), max(0, min(patch, w - x0))
        if ys and xs:
            raw[:, :, :ys, :xs] = tile[:, order, y0:y0 + ys, x0:x0 + xs].astype(F32)
        out[i, :t] = ((raw - m) / s).astype(F32)
    return out


def softmax_classes(logits: np.ndarray) -> np.ndarray:
    """``torch.nn.Softmax(dim=1)`` in fp32 (prediction.py:318): exp(x - max) / sum."""
    x = logits.astype(F32)
    e = np.exp(x - x.max(axis=1, keepdims=True)).astype(F32)
    return (e / e.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)


def classmap_from_logits(logits: np.ndarray, h: int, w: int, patch: int = 128, patch_begin: int = 0,
                         classmap: Optional[np.ndarray] = None, proba: Optional[np.ndarray] = None,
                         grid: Optional[Tuple[int, int]] = None):
    """logits[P, K, patch, patch] of the patches patch_begin .. patch_begin + P - 1 (row-major over the padded tile)
    -> (classmap[h, w] uint8, proba[K, h, w] float32): softmax over K, FIRST maximum (``pred_.max(dim=1)[1]``,
    prediction.py:320), patches put back (prediction.py:329-330) and cropped to the tile (prediction.py:332-333)."""
    p, k = logits.shape[:2]
    gh, gw = patch_grid(h, w, patch) if grid is None else grid
    classmap = np.zeros((h, w), dtype=np.uint8) if classmap is None else classmap
    proba = np.zeros((k, h, w), dtype=F32) if proba is None else proba
    pr = softmax_classes(logits)
    top = pr.argmax(axis=1).astype(np.uint8)  # numpy's argmax returns the first maximum, like torch.max
    for i in range(p):
        ph, pw = divmod(patch_begin + i, gw)
        y0, x0 = ph * patch, pw * patch
        ys, xs = min(patch, h - y0), min(patch, w - x0)
        if ys <= 0 or xs <= 0:
            continue
        classmap[y0:y0 + ys, x0:x0 + xs] = top[i, :ys, :xs]
        proba[:, y0:y0 + ys, x0:x0 + xs] = pr[i, :, :ys, :xs]
    return classmap, proba
