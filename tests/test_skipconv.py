"""Aggregation fused with the decoder's skip convolution: oracle vs the reference's
``UpConvBlock.skip_conv(TemporalAggregator(...))`` (fixtures from tests/golden/make_skipconv_golden.py) and the CUDA
kernel (``c2s_agg_skipconv_forward`` through ``TemporalAggregator.forward_skip_conv``) vs both."""
import numpy as np
import pytest
import torch
from torch import nn

from golden_util import fixture_names, load, rel_err
from oracle import aggregate_skip_conv, temporal_aggregator

TOL_BF16 = 1e-2


@pytest.mark.parametrize("name", fixture_names(["skipconv_"]))
def test_oracle_matches_reference(name):
    cfg, inp, params, outs = load(name)
    skip = temporal_aggregator(inp["x"], inp["pad_mask"], inp["attn"], cfg["mode"])
    assert rel_err(skip, outs["skip"]) < 2e-6
    out = aggregate_skip_conv(inp["x"], inp["pad_mask"], inp["attn"], params, eps=cfg["eps"])
    assert out.shape == outs["out"].shape
    assert rel_err(out, outs["out"]) < 1e-5


def _skip_conv_module(params, device):
    m = nn.Sequential(nn.Conv2d(64, 64, 1), nn.BatchNorm2d(64), nn.ReLU())
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in params.items()}, strict=True)
    assert not missing and not unexpected  # the reference block's state_dict loads as it is
    return m.to(device).eval()


def _random_block(rng):
    p = {"0.weight": (rng.standard_normal((64, 64, 1, 1)) / 8).astype(np.float32),
         "0.bias": (0.1 * rng.standard_normal(64)).astype(np.float32),
         "1.weight": (1.0 + 0.3 * rng.standard_normal(64)).astype(np.float32),
         "1.bias": (0.2 * rng.standard_normal(64)).astype(np.float32),
         "1.running_mean": (0.3 * rng.standard_normal(64)).astype(np.float32),
         "1.running_var": rng.uniform(0.5, 2.0, 64).astype(np.float32),
         "1.num_batches_tracked": np.array(3, dtype=np.int64)}
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("name", fixture_names(["skipconv_"]))
def test_cuda_matches_reference(name):
    import crop2seg_b200 as c2s
    from crop2seg_b200 import _lib
    from c2s_testlib import to_dev
    cfg, inp, params, outs = load(name)
    conv = _skip_conv_module(params, "cuda")
    agg = c2s.TemporalAggregator(mode="att_group")
    with torch.no_grad():
        out = agg.forward_skip_conv(to_dev(inp["x"], dtype=torch.bfloat16), to_dev(inp["pad_mask"]), to_dev(inp["attn"]), conv)
    assert _lib.last_kernel().startswith("agg_skipconv<")
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == outs["out"].shape
    assert rel_err(out.float().cpu().numpy(), outs["out"]) < TOL_BF16


@pytest.mark.gpu
@pytest.mark.parametrize("res,ares,b,t,lengths", [(64, 8, 2, 9, [9, 5]), (32, 16, 3, 61, [61, 27, 0]), (128, 16, 1, 6, [6]),
                                                  (16, 8, 2, 1, [1, 1])])
def test_cuda_matches_oracle_and_unfused_path(res, ares, b, t, lengths):
    """x2 / x4 / x8, series without a valid frame, single-frame series; also against our own unfused kernels + torch."""
    import crop2seg_b200 as c2s
    from c2s_testlib import bf16_round, random_attention, synth_inputs, to_dev
    rng = np.random.RandomState(res + t)
    x, _, pad = synth_inputs(rng, b, t, 64, res, res, lengths)
    x = bf16_round(x)
    attn = random_attention(rng, 16, pad, ares, ares)
    params = _random_block(rng)
    conv = _skip_conv_module(params, "cuda")
    agg = c2s.TemporalAggregator(mode="att_group")
    xd, pd, ad = to_dev(x, dtype=torch.bfloat16), to_dev(pad), to_dev(attn)
    with torch.no_grad():
        fused = agg.forward_skip_conv(xd, pd, ad, conv)
        unfused = conv(agg(xd, pad_mask=pd, attn_mask=ad).float())
    ref = aggregate_skip_conv(x, pad, attn, params, round_skip=bf16_round)
    assert rel_err(fused.float().cpu().numpy(), ref) < TOL_BF16
    assert rel_err(fused.float().cpu().numpy(), unfused.cpu().numpy()) < TOL_BF16
    if 0 in lengths:  # an all-padded series aggregates to zero: the output is relu(shift) for every pixel
        i = lengths.index(0)
        row = fused[i].float().cpu().numpy()
        assert np.abs(row - row[:, :1, :1]).max() == 0.0


@pytest.mark.gpu
def test_unsupported_shapes_fail_loudly():
    import crop2seg_b200 as c2s
    from crop2seg_b200._lib import C2SError
    rng = np.random.RandomState(0)
    conv = _skip_conv_module(_random_block(rng), "cuda")
    agg = c2s.TemporalAggregator(mode="att_group")
    x = torch.zeros((1, 3, 64, 16, 16), device="cuda")  # fp32 features: no fused kernel
    attn = torch.full((16, 1, 3, 8, 8), 1 / 3, device="cuda")
    with pytest.raises(C2SError):
        agg.forward_skip_conv(x, None, attn, conv)
    conv.train()
    with pytest.raises(RuntimeError):
        agg.forward_skip_conv(x.bfloat16(), None, attn, conv)
    with pytest.raises(NotImplementedError):
        c2s.TemporalAggregator(mode="mean").forward_skip_conv(x.bfloat16(), None, attn, conv.eval())
