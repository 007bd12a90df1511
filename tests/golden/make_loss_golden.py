#!/usr/bin/env python
"""Golden vectors for the loss side of the training step, produced by the reference's OWN
code: ``get_dilated`` (src/learning/utils.py:198-222) and the boundary-label statement (:285), ``FocalCELoss``
(src/learning/focal_loss.py) with autograd, and ``nn.CrossEntropyLoss(weight=weights)`` as train.py:462-467 builds it.
Runs only in the build container; the absent GIS / UI packages that src.learning.utils imports are replaced by empty
stand-ins exactly as in make_tile_golden.py.

    python tests/golden/make_loss_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_tile_golden as mtg  # noqa: E402

REF = mtg.REF


def synth_labels(seed, b, h, w, k, blocks=4):
    """Labels U{0..k-1} in blocks x blocks squares, a few single-pixel islands."""
    rng = np.random.RandomState(seed)
    coarse = rng.randint(0, k, size=(b, (h + blocks - 1) // blocks, (w + blocks - 1) // blocks))
    y = np.repeat(np.repeat(coarse, blocks, axis=1), blocks, axis=2)[:, :h, :w].astype(np.int64)
    for _ in range(6):
        y[rng.randint(b), rng.randint(h), rng.randint(w)] = rng.randint(k)
    return y


def synth_scores(seed, b, k, h, w):
    return (np.random.RandomState(seed).standard_normal((b, k, h, w)) * 2.5).astype(np.float32)


def main():
    mtg._stub_missing_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        for _ in range(50):
            try:
                from src.learning.utils import get_dilated
                from src.learning.focal_loss import FocalCELoss
                break
            except ModuleNotFoundError as e:
                top = (e.name or "").split(".")[0]
                if not top or top == "src" or top in mtg.STUB_TOPLEVEL:
                    raise
                mtg.STUB_TOPLEVEL.add(top)
                for name in [n for n in sys.modules if n.startswith("src.")]:
                    del sys.modules[name]
    finally:
        os.chdir(cwd)
    arrays, cfg = {}, {}
    b, k, h, w = 3, 15, 20, 28
    y = synth_labels(7101, b, h, w, k)
    yt = torch.from_numpy(y)
    for conn in (4, 8):
        dil = get_dilated(yt, k, "cpu", conn)
        arrays[f"boundary{conn}"] = torch.where(dil.sum(1) > 1, 1, 0).numpy().astype(np.int64)  # utils.py:285
    cfg["labels"] = dict(seed=7101, B=b, H=h, W=w, K=k)
    # main head: nn.CrossEntropyLoss(weight) with weights[ignore_index = -1] = 0 (train.py:462-467), plain and smoothed
    scores = synth_scores(7102, b, k, h, w)
    weights = torch.ones(k)
    weights[-1] = 0
    for name, eps in (("ce", 0.0), ("ce_smooth", 0.1)):
        z = torch.from_numpy(scores).requires_grad_(True)
        loss = torch.nn.CrossEntropyLoss(weight=weights, label_smoothing=eps)(z, yt)
        (loss * 1.7).backward()
        arrays[f"{name}::loss"] = np.float32(loss.item())
        arrays[f"{name}::grad"] = z.grad.numpy()
    cfg["ce"] = dict(seed=7102, grad_scale=1.7, label_smoothing=[0.0, 0.1])
    # boundary head: FocalCELoss(gamma=2.0) on [B, 2, H, W] against y_b (utils.py:269, 318)
    yb = torch.from_numpy(arrays["boundary4"])
    sb = synth_scores(7103, b, 2, h, w)
    z = torch.from_numpy(sb).requires_grad_(True)
    loss = FocalCELoss(gamma=2.0)(z, yb)
    loss.backward()
    arrays["focal::loss"], arrays["focal::grad"] = np.float32(loss.item()), z.grad.numpy()
    # the other constructor arguments: weight, ignore_index, sum reduction, 15 classes
    yi = y.copy()
    yi[0, :3] = 4
    z = torch.from_numpy(scores).requires_grad_(True)
    wt = torch.linspace(0.5, 2.0, k)
    loss = FocalCELoss(gamma=1.5, size_average=False, ignore_index=4, weight=wt)(z, torch.from_numpy(yi))
    loss.backward()
    arrays["focal_w::loss"], arrays["focal_w::grad"], arrays["focal_w::target"] = np.float32(loss.item()), z.grad.numpy(), yi
    cfg["focal"] = dict(seed_scores_b=7103, gamma=2.0, other=dict(gamma=1.5, size_average=False, ignore_index=4))
    path = os.path.join(HERE, "loss_side.npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), **arrays)
    print(f"loss_side: {os.path.getsize(path) / 1024:.1f} KiB; boundary pixels (4-conn): {int(arrays['boundary4'].sum())} of {b * h * w}")


if __name__ == "__main__":
    main()
