"""Loss side of the training step on the device.

    boundary_target(y)       ``torch.where(get_dilated(y, K, device, 4).sum(1) > 1, 1, 0)``  (learning/utils.py:198-222, 283-285)
    CrossEntropyLoss         ``nn.CrossEntropyLoss(weight=..., label_smoothing=...)`` on [B,K,H,W]   (train.py:462-467)
    FocalCELoss              ``src.learning.focal_loss.FocalCELoss`` (same constructor), used with gamma 2 for the boundary
                             head (learning/utils.py:269, 318)

Forward and backward are kernels of ``libcrop2seg_b200.so`` (``c2s_seg_loss_forward/backward``, ``c2s_boundary_target``);
there is no fallback.  The reductions are deterministic.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import nn

from . import _lib
from .ops import _dtype_code, _require_cuda, _stream


def boundary_target(target: torch.Tensor, n_classes: Optional[int] = None, connectivity: int = 4) -> torch.Tensor:
    """int64 [B,H,W]: 1 on pixels whose (zero-padded) ``connectivity`` neighbourhood holds more than one class, else 0.
    ``n_classes`` is accepted for symmetry with ``get_dilated`` (labels must lie in [0, n_classes) there: F.one_hot)."""
    _require_cuda(target, "target")
    if target.dim() != 3:
        raise RuntimeError(f"crop2seg_b200: target must be [B,H,W], got {tuple(target.shape)}")
    y = target.long().contiguous()
    out = torch.empty_like(y)
    b, h, w = y.shape
    with torch.cuda.device(y.device):
        status = _lib.load().c2s_boundary_target(y.data_ptr(), b, h, w, int(connectivity), out.data_ptr(), _stream(y.device))
    _lib.check(status, "c2s_boundary_target")
    return out


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, target, weight, kind, gamma, size_average, ignore_index, smoothing):
        _require_cuda(scores, "scores")
        if scores.dim() == 2:  # (N, K): FocalCELoss also takes flat rows (focal_loss.py:19)
            scores4, tgt = scores.t().contiguous().view(1, scores.shape[1], 1, scores.shape[0]), target.reshape(1, 1, -1)
        elif scores.dim() == 4:
            scores4, tgt = scores.contiguous(), target
        else:
            raise RuntimeError(f"crop2seg_b200: scores must be [B,K,H,W] or [N,K], got {tuple(scores.shape)}")
        b, k, h, w = scores4.shape
        tgt = tgt.long().contiguous()
        if tgt.numel() != b * h * w:
            raise RuntimeError(f"crop2seg_b200: target {tuple(target.shape)} does not match scores {tuple(scores.shape)}")
        wt = None if weight is None else weight.to(device=scores.device, dtype=torch.float32).contiguous()
        if wt is not None and wt.numel() != k:
            raise RuntimeError(f"crop2seg_b200: weight has {wt.numel()} entries for {k} classes")
        lib = _lib.load()
        desc = _lib.LossDesc(B=b, K=k, H=h, W=w, dtype=_dtype_code(scores4, "scores"), kind=kind, ignore_index=int(ignore_index),
                             size_average=int(bool(size_average)), gamma=float(gamma), label_smoothing=float(smoothing))
        ws = torch.empty(lib.c2s_seg_loss_workspace_bytes() // 8 + 1, dtype=torch.float64, device=scores.device)
        loss = torch.empty((), dtype=torch.float32, device=scores.device)
        with torch.cuda.device(scores.device):
            status = lib.c2s_seg_loss_forward(ctypes.byref(desc), scores4.data_ptr(), tgt.data_ptr(),
                                              None if wt is None else wt.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                              ws.numel() * 8, _stream(scores.device))
        _lib.check(status, "c2s_seg_loss_forward")
        ctx.desc, ctx.flat = desc, scores.dim() == 2
        ctx.save_for_backward(scores4, tgt, wt if wt is not None else torch.empty(0, device=scores.device), ws)
        ctx.has_weight = wt is not None
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        scores4, tgt, wt, ws = ctx.saved_tensors
        g = torch.empty_like(scores4)
        gl = grad_loss.to(torch.float32).contiguous()
        with torch.cuda.device(scores4.device):
            status = _lib.load().c2s_seg_loss_backward(ctypes.byref(ctx.desc), scores4.data_ptr(), tgt.data_ptr(),
                                                       wt.data_ptr() if ctx.has_weight else None, ws.data_ptr(),
                                                       gl.data_ptr(), g.data_ptr(), _stream(scores4.device))
        _lib.check(status, "c2s_seg_loss_backward")
        if ctx.flat:
            g = g.view(g.shape[1], -1).t()
        return g, None, None, None, None, None, None, None


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss(weight=weight, label_smoothing=label_smoothing)`` (mean reduction) for [B,K,H,W] scores
    and int64 [B,H,W] targets, as train.py:462-467 builds it (the ignored class has weight 0)."""

    def __init__(self, weight: Optional[torch.Tensor] = None, label_smoothing: float = 0.0):
        super().__init__()
        self.register_buffer("weight", weight)
        self.label_smoothing = label_smoothing

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _SegLoss.apply(preds, target, self.weight, _lib.LOSS_CROSS_ENTROPY, 0.0, True, -100, self.label_smoothing)


class FocalCELoss(nn.Module):
    """Drop-in for ``src.learning.focal_loss.FocalCELoss`` (focal_loss.py:7-44): same constructor, same result."""

    def __init__(self, gamma=1.0, size_average=True, ignore_index: int = -100, weight: Optional[torch.Tensor] = None):
        super().__init__()
        self.gamma, self.size_average, self.ignore_index, self.weight = gamma, size_average, ignore_index, weight

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        return _SegLoss.apply(preds, target, self.weight, _lib.LOSS_FOCAL, self.gamma, self.size_average, self.ignore_index,
                              0.0)
