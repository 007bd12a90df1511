"""Pin the numpy oracle against outputs of the imported reference (tests/golden/*.npz)."""
import numpy as np
import pytest

from oracle import LtaeConfig, ltae4wtae_forward, ltae_forward, temporal_aggregator
from golden_util import fixture_names, load, load_grads, rel_err

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


# ---- training mode: the reference's own dropout realisation and autograd (tests/golden/make_train_golden.py) ----
def _train_case(name):
    cfg, inp, params, outs = load(name)
    kw = dict(cfg["kwargs"])
    if cfg["kind"] != "ltae":
        kw["mlp"] = [kw["d_model"], 1]
    return cfg, _cfg(kw), inp, params, outs


@pytest.mark.parametrize("name", fixture_names(["train_"]))
def test_numpy_oracle_training_mode_matches_reference(name):
    cfg, c, inp, params, outs = _train_case(name)
    b, t, _, h, w = inp["x"].shape
    n = b * h * w
    ak = np.ascontiguousarray(inp["attn_keep"].transpose(0, 1, 3, 4, 2)).reshape(c.n_head, n, 1, t)
    if cfg["kind"] == "ltae":
        mk = np.ascontiguousarray(inp["mlp_keep"].transpose(0, 2, 3, 1)).reshape(n, -1)
        out, attn, (rm, rv) = ltae_forward(c, params, inp["x"], inp["positions"], inp["pad_mask"], training=True,
                                           attn_keep=ak, mlp_keep=mk)
        assert rel_err(out, outs["out"]) < 5e-5  # out_norm over 4-8 channels amplifies fp32 rounding
        assert rel_err(rm, outs["running_mean"]) < TOL and rel_err(rv, outs["running_var"]) < TOL
    else:
        attn = ltae4wtae_forward(c, params, inp["x"], inp["positions"], inp["pad_mask"], attn_keep=ak)
    assert rel_err(attn, outs["attn"]) < TOL
    assert np.all(attn[inp["attn_keep"] == 0] == 0.0)  # dropout acts on the attention that is returned (tae.py:837)


@pytest.mark.parametrize("name", fixture_names(["train_"]))
def test_torch_port_training_mode_and_gradients_match_reference(name):
    """The differentiable torch-CPU port reproduces the reference's training-mode outputs AND its autograd results
    (grad_x, every parameter gradient) on the reference's own dropout realisation: it is the pinned gradient oracle."""
    import torch
    from oracle.torch_port import ltae_forward_torch
    cfg, c, inp, params, outs = _train_case(name)
    grads = load_grads(name)
    P = {k: torch.from_numpy(v).clone() for k, v in params.items()}
    for k in grads:
        if k != "x":
            P[k].requires_grad_(True)
    x = torch.from_numpy(inp["x"]).requires_grad_(True)
    res = ltae_forward_torch(c, P, x, torch.from_numpy(inp["positions"]), torch.from_numpy(inp["pad_mask"]),
                             attn_only=cfg["kind"] != "ltae", training=True, attn_keep=inp["attn_keep"],
                             mlp_keep=inp.get("mlp_keep"))
    if cfg["kind"] == "ltae":
        out, attn, _ = res
        assert rel_err(out.detach().numpy(), outs["out"]) < 5e-5
        loss = (out * torch.from_numpy(inp["w_out"])).sum() + (attn * torch.from_numpy(inp["w_attn"])).sum()
    else:
        attn = res
        loss = (attn * torch.from_numpy(inp["w_attn"])).sum()
    assert rel_err(attn.detach().numpy(), outs["attn"]) < TOL
    loss.backward()
    assert rel_err(x.grad.numpy(), grads["x"]) < 1e-4
    gmax = max(float(np.abs(g).max()) for k, g in grads.items() if k != "x")
    for k, g in grads.items():
        if k == "x":
            continue
        got = P[k].grad.numpy() if P[k].grad is not None else np.zeros_like(g)
        assert float(np.abs(got - g).max()) <= 1e-4 * max(float(np.abs(g).max()), 1e-2 * gmax), k
