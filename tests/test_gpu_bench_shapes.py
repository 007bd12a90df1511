"""Parity at the shapes bench.py measures (BASELINE.json configs[1] and the Time-Unet placement of configs[2]).

B = 64 patches, T = 61, bf16: L-TAE on x[64,61,128,16,16] and TemporalAggregator on x[64,61,64,{32,64,128}^2] --
the 128^2 feature tensor holds 4.09 G elements = 8.2 GB, so element offsets beyond 2^31 and byte offsets beyond 2^32
(and the TMA tensor maps at their real extents) are exercised.  Ragged lengths include 0 (a series without a valid frame) and 61.  Samples 0, the
all-padded one, a middle one and the LAST one go through the oracle on the host.
"""
import numpy as np
import pytest
import torch

import crop2seg_b200 as c2s
from c2s_testlib import hot_path_parity, randomise

pytestmark = pytest.mark.gpu

B, T = 64, 61


def _lengths(seed):
    rng = np.random.RandomState(seed)
    lengths = rng.randint(27, T + 1, size=B)
    lengths[0], lengths[5], lengths[B - 1] = T, 0, 27
    return lengths


def _inputs(lengths, seed, dev):
    rng = np.random.RandomState(seed)
    pos = np.zeros((B, T), dtype=np.int64)
    pad = np.zeros((B, T), dtype=bool)
    for i, L in enumerate(lengths):
        pad[i, L:] = True
        if L:
            gaps = rng.randint(2, 11, size=L)
            gaps[0] = rng.randint(0, 11)
            pos[i, :L] = np.cumsum(gaps)
    return torch.from_numpy(pos).to(dev), torch.from_numpy(pad).to(dev)


def _feat(c, r, pad, gen, dev, dtype=torch.bfloat16):
    x = torch.empty((B, T, c, r, r), dtype=dtype, device=dev)
    for i in range(B):  # per sample: bounds the fp32 temporaries
        v = torch.randn((T, c, r, r), device=dev, generator=gen).clamp_(min=0)
        v[pad[i]] = 0
        x[i] = v.to(dtype)
    return x


def test_utae_step_at_benchmark_shapes_matches_oracle():
    dev = torch.device("cuda")
    lengths = _lengths(11)
    pos, pad = _inputs(lengths, 12, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(13)
    x4 = _feat(128, 16, pad, gen, dev)
    xs = [_feat(64, r, pad, gen, dev) for r in (32, 64, 128)]
    assert xs[2].numel() > 2 ** 31 and xs[2].numel() * 2 > 2 ** 32  # beyond int32 element and uint32 byte offsets
    enc = c2s.LTAE(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256)
    randomise(enc, np.random.RandomState(14))
    enc = enc.to(dev).eval()
    enc.assume_zero_padded = True
    agg = c2s.TemporalAggregator("att_group")
    with torch.no_grad():
        out, attn = enc(x4, batch_positions=pos, pad_mask=pad)
        skips = [agg(x, pad_mask=pad, attn_mask=attn) for x in xs]
    torch.cuda.synchronize()
    errs = hot_path_parity(enc, x4, xs, pos, pad, out, attn, skips, samples=[0, 5, 31, B - 1])
    print("parity at B=64:", errs)
    assert errs["attn"] < 1e-2 and errs["out"] < 1e-2 and max(errs["skips"]) < 1e-2, errs
    assert errs["pad_attention_exactly_zero"]
    s = attn.sum(dim=2)
    assert float((s - 1).abs().max()) < 1e-4
    # the all-padded series: uniform attention (tae.py:831 fills -1e6 everywhere), zero aggregation
    assert float((attn[:, 5] - 1.0 / T).abs().max()) < 1e-6
    assert all(float(sk[5].abs().max()) == 0.0 for sk in skips)


@pytest.mark.parametrize("return_att", [False, True])
def test_timeunet_ltae_at_benchmark_shapes_matches_oracle(return_att):
    """LTAE(in_channels=64, mlp=[256, 64]) on x[B,61,64,128,128] (timeunet.py:155-180): 16384 pixels per patch."""
    from oracle import LtaeConfig, ltae_forward
    from c2s_testlib import oracle_params
    dev = torch.device("cuda")
    lengths = _lengths(21)
    lengths[:] = T  # configs[2]: all series full length ...
    lengths[5], lengths[B - 1] = 0, 27  # ... except the edge cases
    pos, pad = _inputs(lengths, 22, dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(23)
    x = _feat(64, 128, pad, gen, dev)
    enc = c2s.LTAE(in_channels=64, n_head=16, d_k=4, mlp=[256, 64], d_model=256)
    randomise(enc, np.random.RandomState(24))
    enc = enc.to(dev).eval()
    enc.assume_zero_padded = True
    with torch.no_grad():
        out, attn = enc(x, batch_positions=pos, pad_mask=pad, return_att=return_att)
    torch.cuda.synchronize()
    assert (attn is None) == (not return_att)
    cfg = LtaeConfig(in_channels=64, n_head=16, d_k=4, mlp=[256, 64], d_model=256)
    params = oracle_params(enc)
    rows = slice(40, 48)  # 8 image rows = 1024 pixels of the sample: the oracle materialises [N, T, 256] activations
    for b in (0, 5, B - 1):
        xb = x[b:b + 1, :, :, rows].float().cpu().numpy()
        ref_out, ref_attn = ltae_forward(cfg, params, xb, pos[b:b + 1].cpu().numpy(), pad[b:b + 1].cpu().numpy())
        got = out[b:b + 1, :, rows].float().cpu().numpy()
        assert np.abs(got - ref_out).max() / np.abs(ref_out).max() < 1e-2, b
        if return_att:
            a = attn[:, b:b + 1, :, rows].cpu().numpy()
            assert np.abs(a - ref_attn).max() / np.abs(ref_attn).max() < 1e-2, b
            if 0 < lengths[b] < T:
                assert np.all(a[:, 0, lengths[b]:] == 0.0)
