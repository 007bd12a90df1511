#!/usr/bin/env python
"""Golden vectors for the shared convolutional encoder blocks.

Runs ONLY in the build container: imports the unmodified reference ``ConvBlock`` / ``DownConvBlock``
(src/backbones/conv.py:164-200, 238-296) from ``/root/reference`` and stores ``block.smart_forward(x)`` (eval mode, fp32,
CPU) for seeded inputs with one padded frame, together with the block's ``state_dict``.  The inputs are regenerated from
the seed by the tests (``synth_frames``; already rounded to bfloat16 so that the fp32 reference and the bf16 kernels see
identical values) and are not stored.

    python tests/golden/make_conv_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # U-TAE's in_conv (utae.py:128-136) on a strip of the full-resolution width
    "conv_block_in": dict(kind="ConvBlock", kwargs=dict(nkernels=[10, 64, 64], pad_value=0, norm="group"),
                          shape=(1, 3, 10, 12, 128), padded=[(0, 2)], relu_input=False, seed=21),
    # a second ConvBlock whose first layer already has 64 input channels (both layers on the tensor-core kernel)
    "conv_block_64": dict(kind="ConvBlock", kwargs=dict(nkernels=[64, 64], pad_value=0, norm="group"),
                          shape=(2, 2, 64, 9, 128), padded=[(1, 1)], relu_input=True, seed=22),
    # U-TAE's first down block (utae.py:148-160): 4x4 / stride 2, conv1, out + conv2(out)
    "down_block": dict(kind="DownConvBlock", kwargs=dict(d_in=64, d_out=64, k=4, s=2, p=1, pad_value=0, norm="group"),
                       shape=(1, 3, 64, 16, 128), padded=[(0, 1)], relu_input=True, seed=23),
}


def synth_frames(seed, shape, padded, relu_input):
    """x[B, T, C, H, W] float32 holding bfloat16-representable values, zeros on the padded (b, t) frames."""
    rng = np.random.RandomState(seed)
    x = rng.standard_normal(shape).astype(np.float32)
    if relu_input:
        x = np.maximum(x, 0)
    for b, t in padded:
        x[b, t] = 0
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def randomise_block(block, seed):
    """Seeded, non-trivial parameters: default-initialised convolutions, GroupNorm scale / shift away from 1 / 0."""
    rng = np.random.RandomState(seed + 1000)
    with torch.no_grad():
        for m in block.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.copy_(torch.from_numpy((1.0 + 0.3 * rng.standard_normal(m.num_channels)).astype(np.float32)))
                m.bias.copy_(torch.from_numpy((0.2 * rng.standard_normal(m.num_channels)).astype(np.float32)))
            elif isinstance(m, torch.nn.Conv2d):
                m.bias.copy_(torch.from_numpy((0.1 * rng.standard_normal(m.out_channels)).astype(np.float32)))


def main():
    sys.path.insert(0, REF)
    from src.backbones import conv as ref_conv

    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        block = getattr(ref_conv, c["kind"])(**c["kwargs"])
        randomise_block(block, c["seed"])
        block.eval()
        x = synth_frames(c["seed"], c["shape"], c["padded"], c["relu_input"])
        with torch.no_grad():
            out = block.smart_forward(torch.from_numpy(x))
        arrays = {"cfg": np.array(json.dumps({k: c[k] for k in ("kind", "kwargs", "shape", "padded", "relu_input", "seed")})),
                  "out": out.numpy()}
        for k, v in block.state_dict().items():
            arrays["param::" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, tuple(out.shape), os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
