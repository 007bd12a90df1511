"""Edges of the tile inference pipeline on the device.

What the reference does on the host, patch by patch, before and after its model:

    DatasetCreator._patchify            src/helpers/dataset_creator.py:347-388   zero-pad + cut into 128 x 128 patches
    S2TSCZCropDataset.__getitem__       src/datasets/s2_ts_cz_crop.py:374,393-398 channel reorder, (d - mean) / std
    pad_collate                         src/utils.py:14-66                       frames behind T = pad_value
    generate_prediction (after model)   src/webapp/prediction.py:316-333         softmax, first maximum, every patch
                                                                                 moved to the host, un-patchify, crop

Here the RAW tile (2-byte reflectances) crosses PCIe once, ``patchify`` cuts normalised model inputs out of it on the
device, and ``ClassMap`` keeps the class map (uint8) and the probabilities on the device until the tile is done.
Both are kernels of ``libcrop2seg_b200.so`` (``c2s_tile_patchify`` / ``c2s_tile_classmap``); there is no fallback.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .ops import _require_cuda, _stream

CHANNELS_LIKE_PASTIS = (2, 1, 0, 4, 5, 6, 3, 7, 8, 9)  # s2_ts_cz_crop.py:248
_RAW = {torch.int16: _lib.RAW_I16, torch.uint16: _lib.RAW_U16, torch.float32: _lib.RAW_F32}
_OUT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def patch_grid(h: int, w: int, patch: int = 128, webapp_padding: bool = False) -> Tuple[int, int]:
    """Patch rows / columns of the zero-padded tile.  ``webapp_padding``: the reference's inference path always pads
    by 182 pixels (dataset_creator.py:386: 1098 -> 1280, i.e. 10 x 10 patches whose last row and column are padding
    only); otherwise just enough patches to cover the tile (10980 -> 11008: 86 x 86)."""
    if webapp_padding:
        return (h + 182) // patch, (w + 182) // patch
    return -(-h // patch), -(-w // patch)


def _desc(t, t_pad, c, h, w, patch, grid, begin, count, src, dst, pad_value) -> _lib.TileDesc:
    return _lib.TileDesc(T=t, T_pad=t_pad, C=c, H=h, W=w, patch=patch, grid_h=grid[0], grid_w=grid[1], patch_begin=begin,
                         patch_count=count, src_dtype=src, dst_dtype=dst, pad_value=float(pad_value))


class TilePatchifier:
    """Cuts normalised model inputs out of a raw tile that lives on the device.

    ``tile``: [T, C, H, W] int16 / uint16 / float32 raw values.  ``mean`` / ``std``: per channel, in the REORDERED
    channel order, as ``prediction.py:244-249`` passes them to the dataset.  ``patches(begin, count)`` returns
    [count, T_pad, C, patch, patch] -- the tensors ``pad_collate`` would have built for those patches."""

    def __init__(self, tile: torch.Tensor, mean: Sequence[float], std: Sequence[float],
                 channels_order: Sequence[int] = CHANNELS_LIKE_PASTIS, patch: int = 128, t_pad: Optional[int] = None,
                 pad_value: float = 0.0, grid: Optional[Tuple[int, int]] = None, dtype: torch.dtype = torch.float32):
        if tile.dim() != 4:
            raise RuntimeError(f"crop2seg_b200: tile must be [T,C,H,W], got {tuple(tile.shape)}")
        _require_cuda(tile, "tile")
        if tile.dtype not in _RAW:
            raise RuntimeError(f"crop2seg_b200: raw tile dtype {tile.dtype}; supported: int16, uint16, float32")
        if dtype not in _OUT:
            raise RuntimeError(f"crop2seg_b200: patch dtype {dtype}; supported: float32, bfloat16")
        self.tile = tile.contiguous()
        t, c, h, w = self.tile.shape
        if len(mean) != c or len(std) != c or len(channels_order) != c:
            raise RuntimeError("crop2seg_b200: mean / std / channels_order need one entry per channel")
        if sorted(int(v) for v in channels_order) != list(range(c)):
            raise RuntimeError("crop2seg_b200: channels_order is not a permutation of the channels")
        dev = tile.device
        self.mean = torch.as_tensor(list(mean), dtype=torch.float64).to(torch.float32).to(dev)
        self.std = torch.as_tensor(list(std), dtype=torch.float64).to(torch.float32).to(dev)
        self.order = torch.as_tensor([int(v) for v in channels_order], dtype=torch.int32, device=dev)
        self.patch, self.t_pad, self.pad_value, self.dtype = patch, t if t_pad is None else int(t_pad), pad_value, dtype
        self.grid = patch_grid(h, w, patch) if grid is None else (int(grid[0]), int(grid[1]))
        self.n_patches = self.grid[0] * self.grid[1]

    def patches(self, begin: int = 0, count: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        count = self.n_patches - begin if count is None else count
        t, c, h, w = self.tile.shape
        shape = (count, self.t_pad, c, self.patch, self.patch)
        if out is None:
            out = torch.empty(shape, dtype=self.dtype, device=self.tile.device)
        elif tuple(out.shape) != shape or out.dtype != self.dtype or not out.is_contiguous():
            raise RuntimeError(f"crop2seg_b200: out must be contiguous {shape} {self.dtype}")
        d = _desc(t, self.t_pad, c, h, w, self.patch, self.grid, begin, count, _RAW[self.tile.dtype], _OUT[self.dtype],
                  self.pad_value)
        with torch.cuda.device(self.tile.device):
            status = _lib.load().c2s_tile_patchify(ctypes.byref(d), self.tile.data_ptr(), self.order.data_ptr(),
                                                   self.mean.data_ptr(), self.std.data_ptr(), out.data_ptr(),
                                                   _stream(self.tile.device))
        _lib.check(status, "c2s_tile_patchify")
        return out


class ClassMap:
    """Device-resident result of a tile: ``classmap`` uint8 [H, W] and (optionally) ``proba`` float32 [K, H, W].

    ``put(logits, begin)`` takes the model outputs [P, K, patch, patch] of the patches begin .. begin + P - 1 (row-major
    over the padded tile) and writes softmax probabilities and the FIRST maximum (``pred_.max(dim=1)[1]``,
    prediction.py:318-320) at their place in the tile, cropped to H x W (prediction.py:329-333)."""

    def __init__(self, h: int, w: int, n_classes: int, device, patch: int = 128, grid: Optional[Tuple[int, int]] = None,
                 with_proba: bool = True):
        self.h, self.w, self.k, self.patch = int(h), int(w), int(n_classes), patch
        self.grid = patch_grid(h, w, patch) if grid is None else (int(grid[0]), int(grid[1]))
        self.classmap = torch.zeros((h, w), dtype=torch.uint8, device=device)
        self.proba = torch.zeros((n_classes, h, w), dtype=torch.float32, device=device) if with_proba else None

    def put(self, logits: torch.Tensor, begin: int = 0) -> None:
        _require_cuda(logits, "logits")
        if logits.dim() != 4 or logits.shape[1] != self.k or tuple(logits.shape[2:]) != (self.patch, self.patch):
            raise RuntimeError(f"crop2seg_b200: logits must be [P,{self.k},{self.patch},{self.patch}], got {tuple(logits.shape)}")
        if logits.dtype not in _OUT:
            raise RuntimeError(f"crop2seg_b200: logits dtype {logits.dtype}; supported: float32, bfloat16")
        if logits.device != self.classmap.device:
            raise RuntimeError("crop2seg_b200: logits and class map live on different devices")
        logits = logits.contiguous()
        d = _desc(1, 1, 1, self.h, self.w, self.patch, self.grid, begin, logits.shape[0], 0, _OUT[logits.dtype], 0.0)
        with torch.cuda.device(logits.device):
            status = _lib.load().c2s_tile_classmap(ctypes.byref(d), logits.data_ptr(), self.k, self.classmap.data_ptr(),
                                                   None if self.proba is None else self.proba.data_ptr(),
                                                   _stream(logits.device))
        _lib.check(status, "c2s_tile_classmap")
