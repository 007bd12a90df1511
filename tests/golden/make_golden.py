#!/usr/bin/env python
"""Generate golden vectors for the L-TAE / TemporalAggregator hot path.

Runs ONLY in the build container: it imports the unmodified reference modules from
``/root/reference`` (read-only mount), feeds them seeded synthetic inputs and stores
inputs, parameters and outputs as small ``.npz`` fixtures next to this file.  The GPU
box has no ``/root/reference``; tests there read the committed fixtures.

    python tests/golden/make_golden.py            # regenerate every fixture

Each fixture holds: ``cfg`` (json string), ``x``, ``positions``, ``pad_mask``
(absent when None), ``param::<state_dict key>`` arrays and ``out::<name>`` arrays.
"""
import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    from src.backbones.tae import LTAE, LTAE4WTAE  # noqa
    from src.backbones.temporal_aggregator import TemporalAggregator  # noqa
    return LTAE, LTAE4WTAE, TemporalAggregator


def synth_inputs(rng, b, t, c, h, w, lengths, doy=False, abs_rel=False):
    """Feature maps relu(N(0,1)) with padded frames exactly zero; day positions with 0 at pads."""
    x = np.maximum(rng.standard_normal((b, t, c, h, w)).astype(np.float32), 0)
    pad = np.zeros((b, t), dtype=bool)
    pos = np.zeros((b, t), dtype=np.int64)
    for i, L in enumerate(lengths):
        pad[i, L:] = True
        x[i, L:] = 0
        if L == 0:
            continue
        gaps = rng.randint(2, 11, size=L)
        gaps[0] = rng.randint(0, 11)
        pos[i, :L] = np.cumsum(gaps)
    if abs_rel:
        doyv = np.where(pad, 0, (pos + 243) % 365)
        positions = np.stack([pos, doyv], axis=-1)
    elif doy:
        positions = np.where(pad, 0, (pos + 243) % 365)
    else:
        positions = pos
    return x, positions, pad


def randomise(module, rng):
    """Non-trivial parameters everywhere (BN running stats included) so no term is hidden."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            scale = 1.0 / np.sqrt(max(p.shape[-1], 1)) if p.dim() > 1 else 0.5
            vals = rng.standard_normal(tuple(p.shape)).astype(np.float32) * scale
            if name.endswith("norm.weight") or name == "mlp.2.weight":
                vals = 1.0 + 0.3 * vals
            p.copy_(torch.from_numpy(vals))
        for name, buf in module.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(torch.from_numpy(rng.standard_normal(tuple(buf.shape)).astype(np.float32) * 0.3))
            elif name.endswith("running_var"):
                buf.copy_(torch.from_numpy(rng.uniform(0.5, 2.0, tuple(buf.shape)).astype(np.float32)))


def save(name, cfg, arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


LTAE_CASES = [
    # name, kind, module kwargs, (B, T, H, W), lengths, extra
    ("ltae_small", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64), (2, 7, 4, 4), [7, 4], {}),
    ("ltae_utae_shape", "ltae", dict(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256), (2, 9, 2, 3), [9, 5], {}),
    ("ltae_timeunet_shape", "ltae", dict(in_channels=64, n_head=16, d_k=4, mlp=[256, 64], d_model=256), (1, 11, 3, 4), [8], {}),
    ("ltae_t61", "ltae", dict(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256), (2, 61, 2, 2), [61, 27], {}),
    ("ltae_doy", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, use_doy=True), (2, 6, 3, 3), [6, 3], {"doy": True}),
    ("ltae_abs_rel", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, use_abs_rel_enc=True), (2, 6, 3, 3), [5, 6], {"abs_rel": True}),
    ("ltae_add_linear", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, add_linear=True), (2, 6, 3, 3), [6, 2], {}),
    ("ltae_doy_add_linear", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, use_doy=True, add_linear=True), (2, 6, 3, 3), [6, 4], {"doy": True}),
    ("ltae_no_dmodel", "ltae", dict(in_channels=64, n_head=4, d_k=8, mlp=[64, 48], d_model=None), (2, 5, 3, 2), [5, 3], {}),
    ("ltae_no_pe", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, positional_encoding=False), (2, 5, 2, 2), [5, 2], {"no_positions": True}),
    ("ltae_no_mask", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64), (2, 5, 2, 2), [5, 5], {"no_mask": True}),
    ("ltae_all_padded", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64), (3, 5, 2, 2), [5, 0, 1], {}),
    ("ltae_float_positions", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64), (2, 7, 2, 2), [7, 4], {"float_positions": True}),
    ("ltae_two_queries", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64, num_queries=2), (2, 5, 2, 2), [5, 3], {}),
    ("ltae_train_bn", "ltae", dict(in_channels=32, n_head=4, d_k=4, mlp=[64, 32], d_model=64), (2, 7, 4, 4), [7, 4], {"train": True}),
    ("wtae_small", "ltae4wtae", dict(in_channels=32, n_head=4, d_k=4, d_model=64), (2, 7, 4, 4), [7, 4], {}),
    ("wtae_shape", "ltae4wtae", dict(in_channels=128, n_head=16, d_k=4, d_model=256), (2, 9, 2, 3), [9, 5], {}),
    ("wtae_abs_rel", "ltae4wtae", dict(in_channels=32, n_head=4, d_k=4, d_model=64, use_abs_rel_enc=True), (2, 6, 3, 3), [5, 6], {"abs_rel": True}),
]

AGG_CASES = [
    # name, mode, heads, (B, T, C, H, W), (ha, wa), lengths, pass_mask
    ("agg_group_x2", "att_group", 4, (2, 5, 8, 8, 8), (4, 4), [5, 3], True),
    ("agg_group_x4", "att_group", 16, (2, 4, 64, 16, 16), (4, 4), [4, 2], True),
    ("agg_group_x8", "att_group", 16, (1, 3, 64, 32, 32), (4, 4), [3], True),
    ("agg_group_nomask", "att_group", 4, (2, 5, 8, 8, 8), (4, 4), [5, 5], False),
    ("agg_group_mask_nopad", "att_group", 4, (2, 5, 8, 8, 8), (4, 4), [5, 5], True),
    ("agg_group_frac", "att_group", 4, (2, 4, 8, 12, 12), (5, 5), [4, 3], True),
    ("agg_group_rect", "att_group", 4, (2, 4, 8, 12, 20), (4, 6), [4, 1], True),
    ("agg_group_same_res", "att_group", 4, (2, 4, 8, 6, 6), (6, 6), [4, 2], True),
    ("agg_group_pool", "att_group", 4, (2, 4, 8, 4, 4), (8, 8), [4, 3], True),
    ("agg_group_all_padded", "att_group", 4, (2, 4, 8, 8, 8), (4, 4), [4, 0], True),
    ("agg_mean_att", "att_mean", 4, (2, 5, 6, 8, 8), (4, 4), [5, 3], True),
    ("agg_mean_att_nomask", "att_mean", 4, (2, 5, 6, 8, 8), (4, 4), [5, 5], False),
    ("agg_mean", "mean", 4, (2, 5, 6, 8, 8), (4, 4), [5, 3], True),
    ("agg_mean_nomask", "mean", 4, (2, 5, 6, 8, 8), (4, 4), [5, 5], False),
]


def main():
    LTAE, LTAE4WTAE, TemporalAggregator = _import_reference()
    torch.manual_seed(0)
    for idx, (name, kind, kw, (b, t, h, w), lengths, extra) in enumerate(LTAE_CASES):
        rng = np.random.RandomState(1000 + idx)
        module = (LTAE if kind == "ltae" else LTAE4WTAE)(**kw)
        randomise(module, rng)
        x, positions, pad = synth_inputs(rng, b, t, kw["in_channels"], h, w, lengths,
                                         doy=extra.get("doy", False), abs_rel=extra.get("abs_rel", False))
        if extra.get("float_positions"):
            positions = positions.astype(np.float32)
        arrays = {"x": x}
        tx = torch.from_numpy(x)
        tpos = None if extra.get("no_positions") else torch.from_numpy(positions)
        tpad = None if extra.get("no_mask") else torch.from_numpy(pad)
        if tpos is not None:
            arrays["positions"] = positions
        if tpad is not None:
            arrays["pad_mask"] = pad
        for k, v in module.state_dict().items():
            arrays["param::" + k] = v.detach().numpy().copy()
        pe = getattr(module, "positional_encoder", None)
        if pe is not None and hasattr(pe, "denom"):  # plain attribute, not in the state_dict
            arrays["param::positional_encoder.denom"] = pe.denom.numpy().copy()
        if extra.get("train"):
            module.train()
            module.mlp[5].p = 0.0  # dropout cannot be reproduced (Philox); BN batch statistics can
            module.attention_head.attention.dropout.p = 0.0
            out, attn = module(tx, batch_positions=tpos, pad_mask=tpad)
            arrays["out::out"] = out.detach().contiguous().numpy()
            arrays["out::attn"] = attn.detach().contiguous().numpy()
            arrays["out::running_mean"] = module.mlp[2].running_mean.numpy().copy()
            arrays["out::running_var"] = module.mlp[2].running_var.numpy().copy()
        else:
            module.eval()
            with torch.no_grad():
                res = module(tx, batch_positions=tpos, pad_mask=tpad)
            if kind == "ltae":
                arrays["out::out"] = res[0].contiguous().numpy()
                arrays["out::attn"] = res[1].contiguous().numpy()
            else:
                arrays["out::attn"] = res.contiguous().numpy()
        cfg = dict(kind=kind, kwargs=kw, train=bool(extra.get("train", False)))
        save(name, cfg, arrays)

    for idx, (name, mode, heads, (b, t, c, h, w), (ha, wa), lengths, pass_mask) in enumerate(AGG_CASES):
        rng = np.random.RandomState(2000 + idx)
        x, _, pad = synth_inputs(rng, b, t, c, h, w, lengths)
        logits = rng.standard_normal((heads, b, t, ha, wa)).astype(np.float32)
        logits = np.where(pad[None, :, :, None, None], -1e6, logits)
        e = np.exp(logits - logits.max(axis=2, keepdims=True))
        attn = (e / e.sum(axis=2, keepdims=True)).astype(np.float32)
        agg = TemporalAggregator(mode=mode)
        with torch.no_grad():
            out = agg(torch.from_numpy(x), pad_mask=torch.from_numpy(pad) if pass_mask else None,
                      attn_mask=torch.from_numpy(attn))
        arrays = {"x": x, "attn": attn, "out::out": out.contiguous().numpy()}
        if pass_mask:
            arrays["pad_mask"] = pad
        save(name, dict(kind="aggregator", mode=mode, heads=heads), arrays)


if __name__ == "__main__":
    main()
