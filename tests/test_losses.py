"""Loss side of the training step.

CPU: ``oracle/loss_oracle.py`` against what the reference's own ``get_dilated`` / ``FocalCELoss`` and torch's
``nn.CrossEntropyLoss`` (as train.py builds it) produced, losses AND gradients (tests/golden/make_loss_golden.py).
GPU (``-m gpu``): ``c2s_boundary_target`` bit for bit, ``c2s_seg_loss_forward/backward`` within 1e-5 (fp32).
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_loss_golden import synth_labels, synth_scores  # noqa: E402  (seeded generators; no reference import)
from oracle.loss_oracle import boundary_target, cross_entropy, focal_ce  # noqa: E402
from golden_util import rel_err  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_side.npz")


def _gold():
    z = np.load(GOLD, allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    c = cfg["labels"]
    y = synth_labels(c["seed"], c["B"], c["H"], c["W"], c["K"])
    scores = synth_scores(cfg["ce"]["seed"], c["B"], c["K"], c["H"], c["W"])
    sb = synth_scores(cfg["focal"]["seed_scores_b"], c["B"], 2, c["H"], c["W"])
    return cfg, z, y, scores, sb


def test_oracle_boundary_labels_equal_get_dilated():
    cfg, z, y, _, _ = _gold()
    for conn in (4, 8):
        assert np.array_equal(boundary_target(y, cfg["labels"]["K"], conn), z[f"boundary{conn}"])
    assert 0 < z["boundary4"].sum() < z["boundary8"].sum()


def test_oracle_losses_equal_the_reference_modules():
    cfg, z, y, scores, sb = _gold()
    w = np.ones(cfg["labels"]["K"])
    w[-1] = 0
    for name, eps in (("ce", 0.0), ("ce_smooth", 0.1)):
        loss, g = cross_entropy(scores, y, w, eps)
        assert abs(loss - z[f"{name}::loss"]) < 2e-6 * abs(loss)
        assert rel_err(g * cfg["ce"]["grad_scale"], z[f"{name}::grad"]) < 2e-6
    loss, g = focal_ce(sb, z["boundary4"], gamma=2.0)
    assert abs(loss - z["focal::loss"]) < 2e-6 * abs(loss) and rel_err(g, z["focal::grad"]) < 2e-6
    o = cfg["focal"]["other"]
    loss, g = focal_ce(scores, z["focal_w::target"], gamma=o["gamma"], size_average=o["size_average"],
                       ignore_index=o["ignore_index"], weight=np.linspace(0.5, 2.0, cfg["labels"]["K"]))
    assert abs(loss - z["focal_w::loss"]) < 2e-6 * abs(loss) and rel_err(g, z["focal_w::grad"]) < 2e-6


gpu = pytest.mark.gpu


@gpu
def test_gpu_boundary_labels_equal_get_dilated():
    import crop2seg_b200 as c2s
    cfg, z, y, _, _ = _gold()
    for conn in (4, 8):
        got = c2s.boundary_target(torch.from_numpy(y).cuda(), cfg["labels"]["K"], conn)
        assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), z[f"boundary{conn}"])
    rng = np.random.RandomState(3)  # sizes that are no multiple of anything, a single row, a single pixel
    for shape in ((2, 1, 37), (1, 129, 3), (1, 1, 1), (16, 128, 128)):
        yy = rng.randint(0, 15, size=shape).astype(np.int64)
        assert np.array_equal(c2s.boundary_target(torch.from_numpy(yy).cuda()).cpu().numpy(), boundary_target(yy, 15))


@gpu
def test_gpu_losses_equal_the_reference_modules():
    import crop2seg_b200 as c2s
    cfg, z, y, scores, sb = _gold()
    k = cfg["labels"]["K"]
    w = torch.ones(k)
    w[-1] = 0
    yt = torch.from_numpy(y).cuda()
    for name, eps in (("ce", 0.0), ("ce_smooth", 0.1)):
        s = torch.from_numpy(scores).cuda().requires_grad_(True)
        loss = c2s.CrossEntropyLoss(weight=w, label_smoothing=eps).cuda()(s, yt)
        (loss * cfg["ce"]["grad_scale"]).backward()
        assert abs(loss.item() - z[f"{name}::loss"]) < 1e-5 * abs(z[f"{name}::loss"])
        assert rel_err(s.grad.cpu().numpy(), z[f"{name}::grad"]) < 1e-5
    s = torch.from_numpy(sb).cuda().requires_grad_(True)
    loss = c2s.FocalCELoss(gamma=2.0)(s, torch.from_numpy(z["boundary4"]).cuda())
    loss.backward()
    assert abs(loss.item() - z["focal::loss"]) < 1e-5 * abs(z["focal::loss"])
    assert rel_err(s.grad.cpu().numpy(), z["focal::grad"]) < 1e-5
    o = cfg["focal"]["other"]
    s = torch.from_numpy(scores).cuda().requires_grad_(True)
    loss = c2s.FocalCELoss(gamma=o["gamma"], size_average=o["size_average"], ignore_index=o["ignore_index"],
                           weight=torch.linspace(0.5, 2.0, k))(s, torch.from_numpy(z["focal_w::target"]).cuda())
    loss.backward()
    assert abs(loss.item() - z["focal_w::loss"]) < 1e-5 * abs(z["focal_w::loss"])
    assert rel_err(s.grad.cpu().numpy(), z["focal_w::grad"]) < 1e-5


@gpu
def test_gpu_losses_at_the_training_shape_flat_rows_bf16_and_determinism():
    """B=16, 15 classes, 128 x 128 (BASELINE configs[3]) against the oracle; [N, K] rows; bf16 scores; two runs agree
    bit for bit (fixed-order reduction); ignored and out-of-range labels get exactly zero gradient."""
    import crop2seg_b200 as c2s
    rng = np.random.RandomState(9)
    b, k, h, w = 16, 15, 128, 128
    y = synth_labels(11, b, h, w, k, blocks=8)
    scores = (rng.standard_normal((b, k, h, w)) * 2).astype(np.float32)
    wt = np.ones(k)
    wt[-1] = 0
    ref_loss, ref_g = cross_entropy(scores, y, wt)
    s = torch.from_numpy(scores).cuda().requires_grad_(True)
    crit = c2s.CrossEntropyLoss(weight=torch.from_numpy(wt).float()).cuda()
    l1 = crit(s, torch.from_numpy(y).cuda())
    l1.backward()
    l2 = crit(s.detach(), torch.from_numpy(y).cuda())
    assert torch.equal(l1.detach(), l2)
    assert abs(l1.item() - ref_loss) < 1e-5 * abs(ref_loss) and rel_err(s.grad.cpu().numpy(), ref_g) < 1e-5
    yb = boundary_target(y, k)
    sb = (rng.standard_normal((b, 2, h, w)) * 2).astype(np.float32)
    ref_loss, ref_g = focal_ce(sb, yb, gamma=2.0)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        s = torch.from_numpy(sb).to(dtype).cuda().requires_grad_(True)
        want_l, want_g = (ref_loss, ref_g) if dtype == torch.float32 else focal_ce(s.detach().float().cpu().numpy(), yb, gamma=2.0)
        loss = c2s.FocalCELoss(gamma=2.0)(s, torch.from_numpy(yb).cuda())
        loss.backward()
        assert abs(loss.item() - want_l) < 1e-5 * abs(want_l)
        assert s.grad.dtype == dtype and rel_err(s.grad.float().cpu().numpy(), want_g) < tol
    rows = torch.from_numpy(scores[0].reshape(k, -1).T.copy()).cuda().requires_grad_(True)  # [N, K]
    yr = y[0].reshape(-1).copy()
    yr[:5] = -100
    ref_loss, ref_g = focal_ce(rows.detach().cpu().numpy(), yr, gamma=2.0)
    loss = c2s.FocalCELoss(gamma=2.0)(rows, torch.from_numpy(yr).cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss) < 1e-5 * abs(ref_loss) and rel_err(rows.grad.cpu().numpy(), ref_g) < 1e-5
    assert float(rows.grad[:5].abs().max()) == 0.0
