#!/usr/bin/env python
"""Training-mode golden vectors at the shipped encoder shapes, gradients included.

Runs ONLY in the build container (imports the unmodified reference from ``/root/reference``).  The reference modules
run in ``train()`` mode with their real dropouts; forward hooks on the two ``nn.Dropout`` modules
(``ScaledDotProductAttention.dropout``, tae.py:819/837, and ``mlp[5]``, tae.py:448) record which elements torch's
generator kept, so that the CUDA path (which takes injected keep masks) can be compared with the reference's own
autograd on exactly the same realisation: outputs, BatchNorm running statistics, grad_x and every parameter gradient.

    python tests/golden/make_train_golden.py

Fixture keys: ``cfg`` (json), ``x``, ``positions``, ``pad_mask``, ``attn_keep`` uint8 [h,B,T,H,W], ``mlp_keep`` uint8
[B,C',H,W] (LTAE only), ``w_out`` / ``w_attn`` (weights of the scalar loss ``sum(out*w_out) + sum(attn*w_attn)``),
``param::*``, ``out::out|attn|running_mean|running_var``, ``grad::x``, ``grad::<parameter name>``.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, randomise, synth_inputs  # noqa: E402

CASES = [
    # name, kind, kwargs, (B, T, H, W), lengths
    ("train_utae", "ltae", dict(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256), (3, 13, 4, 4), [13, 6, 9]),
    ("train_timeunet", "ltae", dict(in_channels=64, n_head=16, d_k=4, mlp=[256, 64], d_model=256), (3, 13, 4, 4), [13, 6, 9]),
    ("train_wtae", "ltae4wtae", dict(in_channels=128, n_head=16, d_k=4, d_model=256), (3, 13, 4, 4), [13, 6, 9]),
]


class Capture:
    """Forward hook: which elements did this nn.Dropout keep?  (Where the input is exactly 0 the answer does not
    matter -- padded frames of the attention, clipped ReLU outputs: value and gradient are 0 either way.)"""

    def __init__(self):
        self.keep = None

    def __call__(self, module, inputs, output):
        x = inputs[0].detach()
        self.keep = ((output.detach() != 0) | (x == 0)).to(torch.uint8)


def main():
    sys.path.insert(0, REF)
    from src.backbones.tae import LTAE, LTAE4WTAE
    for idx, (name, kind, kw, (b, t, h, w), lengths) in enumerate(CASES):
        rng = np.random.RandomState(9000 + idx)
        torch.manual_seed(9000 + idx)
        module = (LTAE if kind == "ltae" else LTAE4WTAE)(**kw)
        randomise(module, rng)
        c = kw["in_channels"]
        x, positions, pad = synth_inputs(rng, b, t, c, h, w, lengths)
        x = x + 0.3 * rng.standard_normal(x.shape).astype(np.float32) * (~pad)[:, :, None, None, None]
        nh = kw["n_head"]
        arrays = {"x": x, "positions": positions, "pad_mask": pad}
        for k, v in module.state_dict().items():
            arrays["param::" + k] = v.detach().numpy().copy()
        arrays["param::positional_encoder.denom"] = module.positional_encoder.denom.numpy().copy()
        module.train()
        cap_attn, cap_mlp = Capture(), Capture()
        module.attention_head.attention.dropout.register_forward_hook(cap_attn)
        if kind == "ltae":
            module.mlp[5].register_forward_hook(cap_mlp)
        tx = torch.from_numpy(x).requires_grad_(True)
        res = module(tx, batch_positions=torch.from_numpy(positions), pad_mask=torch.from_numpy(pad))
        out, attn = res if kind == "ltae" else (None, res)
        w_attn = rng.standard_normal(tuple(attn.shape)).astype(np.float32)
        loss = (attn * torch.from_numpy(w_attn)).sum()
        arrays["w_attn"] = w_attn
        if out is not None:
            w_out = rng.standard_normal(tuple(out.shape)).astype(np.float32)
            loss = loss + (out * torch.from_numpy(w_out)).sum()
            arrays["w_out"] = w_out
        loss.backward()
        # attention dropout sees [h * N, 1, T] with N = (b, y, x) rows (tae.py:764-778) -> [h, B, T, H, W]
        arrays["attn_keep"] = np.ascontiguousarray(
            cap_attn.keep.view(nh, b, h, w, t).permute(0, 1, 4, 2, 3).numpy())
        arrays["out::attn"] = attn.detach().contiguous().numpy()
        if kind == "ltae":
            co = kw["mlp"][-1]
            arrays["mlp_keep"] = np.ascontiguousarray(cap_mlp.keep.view(b, h, w, co).permute(0, 3, 1, 2).numpy())
            arrays["out::out"] = out.detach().contiguous().numpy()
            arrays["out::running_mean"] = module.mlp[2].running_mean.numpy().copy()
            arrays["out::running_var"] = module.mlp[2].running_var.numpy().copy()
        arrays["grad::x"] = tx.grad.numpy().copy()
        for pname, p in module.named_parameters():
            arrays["grad::" + pname] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
        cfg = dict(kind=kind, kwargs=kw, train=True, attn_drop_p=0.1, mlp_drop_p=0.2)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, cfg=json.dumps(cfg), **arrays)
        kept = float(arrays["attn_keep"].mean())
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB, attention keep rate {kept:.3f}")


if __name__ == "__main__":
    main()
