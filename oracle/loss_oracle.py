"""Numpy restatement of the loss side of the reference's training step.  TEST INFRASTRUCTURE ONLY.

    boundary_target     src/learning/utils.py:198-222 (get_dilated) + :283-285 (y_b)
    cross_entropy       nn.CrossEntropyLoss(weight, label_smoothing) as train.py:462-467 builds it (torch semantics:
                        weighted mean over the targets that are not -100; the smoothing term uses the class weights)
    focal_ce            src/learning/focal_loss.py:17-44

Each function returns (loss, gradient with respect to the scores).  Parity status: pinned against the reference's own
``get_dilated`` / ``FocalCELoss`` and torch's ``nn.CrossEntropyLoss`` with autograd (tests/golden/make_loss_golden.py,
tests/test_losses.py).
"""
from __future__ import annotations

import numpy as np

F64 = np.float64


def boundary_target(y: np.ndarray, n_classes: int, connectivity: int = 4) -> np.ndarray:
    """y[B,H,W] int -> int64 [B,H,W]: the as-written algorithm (one-hot, grouped 3x3 convolution, `.bool()`, sum > 1)."""
    b, h, w = y.shape
    one_hot = (y[:, None] == np.arange(n_classes)[None, :, None, None]).astype(np.float32)  # utils.py:220
    padded = np.pad(one_hot, ((0, 0), (0, 0), (1, 1), (1, 1)))                             # padding=(1, 1)
    if connectivity == 8:
        taps = [(dy, dx) for dy in range(3) for dx in range(3)]
    else:
        taps = [(0, 1), (1, 0), (1, 1), (1, 2), (2, 1)]                                    # utils.py:213-216
    conv = sum(padded[:, :, dy:dy + h, dx:dx + w] for dy, dx in taps)
    dilated = (conv != 0).astype(np.int64)                                                  # .bool().long()
    return np.where(dilated.sum(1) > 1, 1, 0).astype(np.int64)                              # utils.py:285


def _log_softmax(z):
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m)
    return z - m - np.log(e.sum(axis=1, keepdims=True))


def cross_entropy(scores: np.ndarray, target: np.ndarray, weight=None, label_smoothing: float = 0.0):
    """scores[B,K,H,W], target[B,H,W] -> (loss, grad[B,K,H,W]) in float64."""
    b, k, h, w = scores.shape
    z = scores.astype(F64).transpose(0, 2, 3, 1).reshape(-1, k)
    y = target.reshape(-1)
    wt = np.ones(k, F64) if weight is None else np.asarray(weight, F64)
    keep = y != -100
    logp = _log_softmax(z)
    p = np.exp(logp)
    yi = np.where(keep, y, 0)
    wy = wt[yi] * keep
    nll = -logp[np.arange(len(y)), yi]
    smooth = -(logp * wt[None, :]).sum(1) * keep
    eps = label_smoothing
    den = wy.sum()
    loss = ((1 - eps) * (wy * nll).sum() + eps / k * smooth.sum()) / den
    onehot = np.zeros_like(z)
    onehot[np.arange(len(y)), yi] = 1.0
    g = (1 - eps) * wy[:, None] * (p - onehot) + eps / k * keep[:, None] * (wt.sum() * p - wt[None, :])
    g = g / den
    return loss, g.reshape(b, h, w, k).transpose(0, 3, 1, 2)


def focal_ce(scores: np.ndarray, target: np.ndarray, gamma: float = 1.0, size_average: bool = True, ignore_index: int = -100,
             weight=None):
    """focal_loss.py:17-44 on scores[B,K,H,W] (or [N,K]) -> (loss, grad) in float64, AS WRITTEN: with a class weight the
    gathered weights keep their [N, 1] shape (:34-35) and broadcast against the [N] focal terms (:36), so the loss is
    the N x N outer product  -(1 - pt_j)^gamma w[y_i] logpt_j  -- its sum is (sum_i w[y_i]) (sum_j focal_j), its mean
    divides by N^2.  (The reference itself only builds FocalCELoss(gamma=2.0), without weights.)"""
    flat = scores.ndim == 2
    if flat:
        z = scores.astype(F64)
    else:
        b, k, h, w = scores.shape
        z = scores.astype(F64).transpose(0, 2, 3, 1).reshape(-1, k)          # :21-22
    k = z.shape[1]
    y = target.reshape(-1)
    keep = y != ignore_index                                                  # :24
    logp = _log_softmax(z)                                                    # :28
    yi = np.where(keep, y, 0)
    logpt = logp[np.arange(len(y)), yi]                                       # :29
    pt = np.exp(logpt)                                                        # :31
    per = -1 * (1 - pt) ** gamma * logpt * keep                               # :38 (and the [N] factor of :36)
    n = keep.sum()
    if weight is None:
        mult = 1.0 / n if size_average else 1.0
    else:
        wsum = (np.asarray(weight, F64)[yi] * keep).sum()                     # the [N, 1] factor of :36
        mult = wsum / (float(n) * float(n)) if size_average else wsum
    loss = per.sum() * mult
    tail = gamma * (1 - pt) ** (gamma - 1) * pt * logpt if gamma != 0 else 0.0
    dl = -((1 - pt) ** gamma - tail)                                          # d per / d logpt
    onehot = np.zeros_like(z)
    onehot[np.arange(len(y)), yi] = 1.0
    g = dl[:, None] * (onehot - np.exp(logp)) * keep[:, None] * mult
    if flat:
        return loss, g
    return loss, g.reshape(b, h, w, k).transpose(0, 3, 1, 2)
