#!/usr/bin/env python
"""Generate the model-level golden fixture: the hot path inside the reference's full U-TAE.

Runs ONLY in the build container (imports the unmodified reference from ``/root/reference``).  It builds
``UTAE(input_dim=10, out_conv=[32, 15])`` with seeded weights (BatchNorm running statistics randomised), runs it in
eval mode on a seeded padded batch and records, with forward hooks, everything that crosses the hot-path boundary:

* what the model hands to ``temporal_encoder`` (the lowest-resolution feature maps, positions, pad mask) and to
  the three ``temporal_aggregator`` calls (the skip feature maps),
* what the reference hot path returns (``out``, ``attn``, the three skip maps),
* the class scores and their per-pixel argmax at the end of the decoder.

The decoder that consumes the hot path's outputs (``up_blocks`` + ``out_conv``, utae.py:223-241) is stored as a
TorchScript trace (``model_utae_decoder.pt``: a serialised graph of torch operators plus this seed's weights, no
reference source), so that the GPU box -- which has no ``/root/reference`` -- can push OUR hot-path outputs through the
reference's own decoder and apply SURVEY.md 8(d)'s gate: per-pixel argmax agreement >= 99.9 %.

    python tests/golden/make_model_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20240611


class Decoder(torch.nn.Module):
    """The part of UTAE.forward after the hot path: three up blocks fed by the skips, then out_conv."""

    def __init__(self, model):
        super().__init__()
        self.up_blocks = model.up_blocks
        self.out_conv = model.out_conv

    def forward(self, out, skip0, skip1, skip2):
        for blk, skip in zip(self.up_blocks, (skip0, skip1, skip2)):
            out = blk(out, skip)
        return self.out_conv(out)


def main():
    sys.path.insert(0, REF)
    from src.backbones.utae import UTAE

    torch.manual_seed(SEED)
    rng = np.random.RandomState(SEED)
    model = UTAE(input_dim=10, out_conv=[32, 15])
    with torch.no_grad():  # BatchNorm running statistics away from (0, 1) so that eval BN is not the identity
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.from_numpy(rng.standard_normal(tuple(buf.shape)).astype(np.float32) * 0.3))
            elif name.endswith("running_var"):
                buf.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, tuple(buf.shape)).astype(np.float32)))
    model.eval()

    b, t, lengths = 2, 5, [5, 3]
    x = rng.standard_normal((b, t, 10, 32, 32)).astype(np.float32)
    pos = np.zeros((b, t), dtype=np.int64)
    for i, n in enumerate(lengths):
        x[i, n:] = 0.0  # pad_value: the model derives its pad mask from all-zero frames (utae.py:201-203)
        gaps = rng.randint(2, 11, size=n)
        gaps[0] = rng.randint(0, 11)
        pos[i, :n] = np.cumsum(gaps)

    seen = {"agg_in": [], "agg_out": []}

    def enc_hook(_m, args, kwargs, output):
        seen["enc_x"] = args[0].detach()
        seen["enc_pos"] = kwargs["batch_positions"].detach()
        seen["enc_pad"] = kwargs["pad_mask"].detach()
        seen["enc_out"], seen["enc_attn"] = (o.detach() for o in output)

    def agg_hook(_m, args, kwargs, output):
        seen["agg_in"].append(args[0].detach())
        seen["agg_out"].append(output.detach())

    h1 = model.temporal_encoder.register_forward_hook(enc_hook, with_kwargs=True)
    h2 = model.temporal_aggregator.register_forward_hook(agg_hook, with_kwargs=True)
    with torch.no_grad():
        logits = model(torch.from_numpy(x), batch_positions=torch.from_numpy(pos))
    h1.remove(), h2.remove()
    assert len(seen["agg_in"]) == 3

    dec = Decoder(model).eval()
    with torch.no_grad():
        traced = torch.jit.trace(dec, (seen["enc_out"], *seen["agg_out"]))
        again = traced(seen["enc_out"], *seen["agg_out"])
    assert torch.equal(again, logits), "the traced decoder must reproduce the model's class scores bit for bit"
    torch.jit.save(traced, os.path.join(HERE, "model_utae_decoder.pt"))

    cfg = {"model": "UTAE(input_dim=10, out_conv=[32, 15])", "seed": SEED, "lengths": lengths,
           "ltae_kwargs": {"in_channels": 128, "n_head": 16, "d_k": 4, "mlp": [256, 128], "d_model": 256},
           "agg_mode": "att_group"}
    arrays = {"cfg": np.array(json.dumps(cfg)), "enc_x": seen["enc_x"].numpy(), "positions": seen["enc_pos"].numpy(),
              "pad_mask": seen["enc_pad"].numpy(),  # what utae.py:201-203 derived from the raw input below
              "raw_input": x}
    for i, a in enumerate(seen["agg_in"]):
        arrays[f"skip_x{i}"] = a.numpy()
    for k, v in model.temporal_encoder.state_dict().items():
        arrays["param::" + k] = v.numpy()
    arrays["out::out"] = seen["enc_out"].numpy()
    arrays["out::attn"] = seen["enc_attn"].numpy()
    for i, a in enumerate(seen["agg_out"]):
        arrays[f"out::skip{i}"] = a.numpy()
    arrays["out::logits"] = logits.numpy()
    arrays["out::argmax"] = logits.argmax(dim=1).numpy().astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "model_utae.npz"), **arrays)
    for f in ("model_utae.npz", "model_utae_decoder.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
