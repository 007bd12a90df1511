#!/usr/bin/env python
"""Golden vectors for the aggregation + skip convolution fusion.

Runs ONLY in the build container: imports the unmodified reference ``TemporalAggregator`` and ``UpConvBlock`` from
``/root/reference`` and stores ``up_block.skip_conv(aggregator(x, pad_mask, attn))`` (eval mode) for seeded inputs.
x is stored already rounded to bfloat16 so that the fp32 reference and the bf16 kernel see identical values.

    python tests/golden/make_skipconv_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    sys.path.insert(0, REF)
    from src.backbones.conv import UpConvBlock
    from src.backbones.temporal_aggregator import TemporalAggregator

    cases = {"skipconv_x2": dict(b=2, t=7, res=16, ares=8, lengths=[7, 4], seed=11),
             "skipconv_x4": dict(b=1, t=4, res=32, ares=8, lengths=[3], seed=12)}
    for name, c in cases.items():
        rng = np.random.RandomState(c["seed"])
        torch.manual_seed(c["seed"])
        blk = UpConvBlock(d_in=64, d_out=64, k=4, s=2, p=1, norm="batch", d_skip=64, padding_mode="reflect")
        with torch.no_grad():
            bn = blk.skip_conv[1]
            bn.weight.copy_(torch.from_numpy(1.0 + 0.3 * rng.standard_normal(64).astype(np.float32)))
            bn.bias.copy_(torch.from_numpy(0.2 * rng.standard_normal(64).astype(np.float32)))
            bn.running_mean.copy_(torch.from_numpy(0.3 * rng.standard_normal(64).astype(np.float32)))
            bn.running_var.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, 64).astype(np.float32)))
        blk.eval()
        b, t, res, ares = c["b"], c["t"], c["res"], c["ares"]
        x = np.maximum(rng.standard_normal((b, t, 64, res, res)).astype(np.float32), 0)
        pad = np.zeros((b, t), dtype=bool)
        for i, n in enumerate(c["lengths"]):
            pad[i, n:] = True
            x[i, n:] = 0
        x = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
        logits = rng.standard_normal((16, b, t, ares, ares)).astype(np.float32)
        logits = np.where(pad[None, :, :, None, None], -1e6, logits)
        e = np.exp(logits - logits.max(axis=2, keepdims=True))
        attn = (e / e.sum(axis=2, keepdims=True)).astype(np.float32)
        with torch.no_grad():
            skip = TemporalAggregator(mode="att_group")(torch.from_numpy(x), pad_mask=torch.from_numpy(pad),
                                                        attn_mask=torch.from_numpy(attn))
            out = blk.skip_conv(skip)
        arrays = {"cfg": np.array(json.dumps({"mode": "att_group", "eps": float(blk.skip_conv[1].eps)})),
                  "x": x, "pad_mask": pad, "attn": attn, "out::skip": skip.numpy(), "out::out": out.numpy()}
        for k, v in blk.skip_conv.state_dict().items():
            arrays["param::" + k] = v.numpy()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, os.path.getsize(os.path.join(HERE, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
