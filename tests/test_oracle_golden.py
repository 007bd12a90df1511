"""Pin the numpy oracle against outputs of the imported reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import LtaeConfig, ltae4wtae_forward, ltae_forward, temporal_aggregator
from golden_util import fixture_names, load, rel_err

TOL = 2e-6  # fp32 restatement vs fp32 reference: a few ulp of the largest element


def _cfg(kwargs):
    kw = dict(kwargs)
    kw.setdefault("mlp", [256, 128])
    return LtaeConfig(**kw)


@pytest.mark.parametrize("name", fixture_names(["ltae_"]))
def test_ltae_matches_reference(name):
    cfg, inp, params, outs = load(name)
    c = _cfg(cfg["kwargs"])
    res = ltae_forward(c, params, inp["x"], inp.get("positions"), inp.get("pad_mask"),
                       training=cfg["train"])
    assert res[0].shape == outs["out"].shape
    assert res[1].shape == outs["attn"].shape
    assert rel_err(res[1], outs["attn"]) < TOL
    assert rel_err(res[0], outs["out"]) < 2e-5  # out_norm divides by tiny group variances
    if cfg["train"]:
        assert rel_err(res[2][0], outs["running_mean"]) < TOL
        assert rel_err(res[2][1], outs["running_var"]) < TOL
    if "pad_mask" in inp:  # attention is exactly zero on padded frames of partly valid series
        pad = inp["pad_mask"]
        some_valid = ~pad.all(axis=1)
        a = res[1]
        if a.ndim == 5:
            for b in np.nonzero(some_valid)[0]:
                assert np.all(a[:, b, pad[b]] == 0.0)


@pytest.mark.parametrize("name", fixture_names(["wtae_"]))
def test_ltae4wtae_matches_reference(name):
    cfg, inp, params, outs = load(name)
    kw = dict(cfg["kwargs"])
    kw["mlp"] = [kw.get("d_model") or kw["in_channels"], 1]
    attn = ltae4wtae_forward(LtaeConfig(**kw), params, inp["x"], inp.get("positions"), inp.get("pad_mask"))
    assert attn.shape == outs["attn"].shape
    assert rel_err(attn, outs["attn"]) < TOL


@pytest.mark.parametrize("name", fixture_names(["agg_"]))
def test_aggregator_matches_reference(name):
    cfg, inp, _, outs = load(name)
    out = temporal_aggregator(inp["x"], inp.get("pad_mask"), inp["attn"], cfg["mode"])
    assert out.shape == outs["out"].shape
    assert rel_err(out, outs["out"]) < TOL


def test_softmax_sums_to_one():
    cfg, inp, params, outs = load("ltae_t61")
    _, attn = ltae_forward(_cfg(cfg["kwargs"]), params, inp["x"], inp["positions"], inp["pad_mask"])
    s = attn.sum(axis=2)
    assert np.all(np.abs(s - 1.0) < 1e-5)


# ---- the torch-CPU port (bench.py's cpu_baseline and a second oracle) is pinned the same way -------------
@pytest.mark.parametrize("name", [n for n in fixture_names(["ltae_", "wtae_"])
                                  if n not in ("ltae_two_queries", "ltae_train_bn")])
def test_torch_port_ltae_matches_reference(name):
    from oracle.torch_port import ltae_forward_torch
    cfg, inp, params, outs = load(name)
    kw = dict(cfg["kwargs"])
    if cfg["kind"] != "ltae":
        kw["mlp"] = [kw.get("d_model") or kw["in_channels"], 1]
    res = ltae_forward_torch(_cfg(kw), params, inp["x"], inp.get("positions"), inp.get("pad_mask"),
                             attn_only=cfg["kind"] != "ltae")
    if cfg["kind"] == "ltae":
        assert rel_err(res[0].numpy(), outs["out"]) < 2e-5
        assert rel_err(res[1].numpy(), outs["attn"]) < TOL
    else:
        assert rel_err(res.numpy(), outs["attn"]) < TOL


@pytest.mark.parametrize("name", fixture_names(["agg_"]))
def test_torch_port_aggregator_matches_reference(name):
    from oracle.torch_port import temporal_aggregator_torch
    cfg, inp, _, outs = load(name)
    out = temporal_aggregator_torch(inp["x"], inp.get("pad_mask"), inp["attn"], cfg["mode"])
    assert rel_err(out.numpy(), outs["out"]) < TOL
