"""Shared convolutional encoder blocks: ``crop2seg_b200.ConvBlock`` / ``DownConvBlock``.

CPU: the ``state_dict`` contract against the reference blocks' own keys (fixtures of tests/golden/make_conv_golden.py),
loud failures for what the inference path does not serve.
GPU (``-m gpu``): the tcgen05 implicit-GEMM convolution against a plain fp32 torch convolution of the same bf16 operands
(bands, whole-frame units, several units per CTA, 10 and 64 input channels), its GroupNorm sums, the normalisation pass,
and the blocks through ``smart_forward`` against what the reference blocks produced.  bf16 features: 1e-2 of max|ref|.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

import crop2seg_b200 as c2s
from crop2seg_b200 import conv as c2s_conv

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_conv_golden import CASES, synth_frames  # noqa: E402  (seeded generator only; no reference import)

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_BF16 = 1e-2


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    return cfg, params, z["out"]


def _block(cfg, params):
    blk = getattr(c2s, cfg["kind"])(**cfg["kwargs"])
    missing, unexpected = blk.load_state_dict(params, strict=True)
    assert not missing and not unexpected
    return blk.eval()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dict_matches_the_reference_block(name):
    cfg, params, _ = _load(name)
    blk = getattr(c2s, cfg["kind"])(**cfg["kwargs"])
    assert list(blk.state_dict().keys()) == list(params.keys())  # same keys, same order
    for k, v in blk.state_dict().items():
        assert tuple(v.shape) == tuple(params[k].shape), k
    _block(cfg, params)


def test_what_is_not_served_fails_loudly():
    with pytest.raises(NotImplementedError):
        c2s.ConvBlock([10, 64], norm="batch")
    with pytest.raises(NotImplementedError):
        c2s.ConvBlock([10, 64], norm="group", conv_type="depthwise_separable")
    blk = c2s.ConvBlock([10, 64], norm="group")  # training mode by default
    with pytest.raises(NotImplementedError):
        blk(torch.zeros(1, 10, 4, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk.eval()(torch.zeros(1, 10, 4, 128))


# ------------------------------------------------------------------------------------------------ GPU
def _torch_conv(x, w, b):
    xp = torch.nn.functional.pad(x.float(), (1, 1, 1, 1), mode="reflect")
    return torch.nn.functional.conv2d(xp, w.to(torch.bfloat16).float(), b)


@pytest.mark.gpu
@pytest.mark.parametrize("frames,c_in,h,w", [(3, 64, 12, 128), (2, 10, 7, 128), (1, 64, 2, 128), (5, 3, 33, 128),
                                             (150, 64, 16, 128), (20, 64, 128, 128),
                                             (3, 64, 64, 64), (7, 64, 5, 64), (2, 10, 9, 64), (300, 64, 64, 64),
                                             (3, 64, 32, 32), (5, 10, 7, 32), (1, 64, 2, 32), (700, 64, 32, 32)])
def test_tensor_core_convolution_matches_torch(frames, c_in, h, w):
    g = torch.Generator(device="cuda").manual_seed(frames * 131 + c_in)
    x = torch.randn((frames, c_in, h, w), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((64, c_in, 3, 3), device="cuda", generator=g) * (1.0 / (3.0 * c_in ** 0.5))
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    conv = torch.nn.Conv2d(c_in, 64, 3, padding=1, padding_mode="reflect")
    assert c2s_conv.conv2d_supported(x, conv)
    y, stats = c2s_conv.conv2d_reflect_forward(x, w, b)
    torch.cuda.synchronize()
    assert c2s.ops._lib.load().c2s_last_kernel().decode() == "conv3x3_reflect<tcgen05>"
    ref = _torch_conv(x, w, b)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, err  # bf16 rounding of the stored output; the products are exact, the sums fp32
    # GroupNorm sums of the fp32 values, per frame and quarter of the channels
    q = ref.view(frames, 4, -1)
    s1, s2 = q.sum(-1), (q * q).sum(-1)
    assert torch.allclose(stats[..., 0], s1, rtol=2e-3, atol=2e-3 * s1.abs().max().item())
    assert torch.allclose(stats[..., 1], s2, rtol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("frames,h,w_in", [(2, 128, 128), (3, 4, 128), (5, 10, 128), (1, 6, 128), (37, 32, 128), (200, 128, 128),
                                            (9, 64, 128), (3, 64, 64), (5, 6, 64), (400, 64, 64), (3, 32, 32), (2, 4, 32),
                                            (900, 32, 32)])
def test_tensor_core_strided_convolution_matches_torch(frames, h, w_in):
    """The strided layer of DownConvBlock (conv.py:252-263): 4x4 / stride 2 / reflect padding 1, 64 -> 64 channels from
    128 / 64 / 32-pixel rows -- parity-split rows, input-stationary products -- against a plain fp32 torch convolution of
    the same bf16 operands: whole frames and bands per unit, the reflected rows at both frame edges, several units per CTA."""
    g = torch.Generator(device="cuda").manual_seed(frames * 17 + h)
    x = torch.randn((frames, 64, h, w_in), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn((64, 64, 4, 4), device="cuda", generator=g) * (1.0 / 32.0)
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    conv = torch.nn.Conv2d(64, 64, 4, stride=2, padding=1, padding_mode="reflect")
    assert c2s_conv.conv2d_supported(x, conv)
    y, stats = c2s_conv.conv2d_reflect_forward(x, w, b, kernel=4, stride=2, padding=1)
    torch.cuda.synchronize()
    assert c2s.ops._lib.load().c2s_last_kernel().decode() == "conv4x4s2_reflect<tcgen05>"
    xp = torch.nn.functional.pad(x.float(), (1, 1, 1, 1), mode="reflect")
    ref = torch.nn.functional.conv2d(xp, w.to(torch.bfloat16).float(), b, stride=2)
    assert y.shape == ref.shape == (frames, 64, h // 2, w_in // 2) and y.dtype == torch.bfloat16
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 6e-3, err
    q = ref.view(frames, 4, -1)
    s1, s2 = q.sum(-1), (q * q).sum(-1)
    assert torch.allclose(stats[..., 0], s1, rtol=2e-3, atol=2e-3 * s1.abs().max().item())
    assert torch.allclose(stats[..., 1], s2, rtol=2e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("residual", [False, True])
def test_group_norm_relu_matches_torch(dtype, residual):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn((3, 64, 8, 24), device="cuda", generator=g) * 2 + 0.5).to(dtype)
    res = torch.randn((3, 64, 8, 24), device="cuda", generator=g).to(dtype) if residual else None
    norm = torch.nn.GroupNorm(4, 64).cuda()
    with torch.no_grad():
        norm.weight.copy_(1 + 0.3 * torch.randn(64, device="cuda", generator=g))
        norm.bias.copy_(0.2 * torch.randn(64, device="cuda", generator=g))
    stats = c2s_conv.group_stats(x, 4)
    out = c2s_conv.group_norm_relu(x, stats, norm, relu=True, residual=res)
    with torch.no_grad():
        ref = torch.relu(norm(x.float()))
        if residual:
            ref = ref + res.float()
    tol = 1e-5 if dtype == torch.float32 else TOL_BF16
    assert _rel(out.float().cpu().numpy(), ref.cpu().numpy()) < tol
    # statistics at a finer granularity (16 sub-groups; the convolution kernel hands over 4) give the same result
    out16 = c2s_conv.group_norm_relu(x, c2s_conv.group_stats(x, 16), norm, relu=True, residual=res)
    assert _rel(out16.float().cpu().numpy(), out.float().cpu().numpy()) < (1e-6 if dtype == torch.float32 else TOL_BF16)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_blocks_match_the_reference(name):
    cfg, params, ref = _load(name)
    blk = _block(cfg, params).cuda()
    x = synth_frames(cfg["seed"], tuple(cfg["shape"]), [tuple(p) for p in cfg["padded"]], cfg["relu_input"])
    with torch.no_grad():
        out = blk.smart_forward(torch.from_numpy(x).cuda().to(torch.bfloat16))
    assert tuple(out.shape) == ref.shape and out.dtype == torch.bfloat16
    got = out.float().cpu().numpy()
    assert _rel(got, ref) < TOL_BF16 * 2, _rel(got, ref)  # two to three bf16 layers deep
    for b, t in cfg["padded"]:  # padded frames come back as pad_value, exactly (temp_shared_block.py:30-40)
        assert np.all(got[b, t] == 0.0)


@pytest.mark.gpu
def test_in_conv_runs_on_the_tensor_core_kernel():
    blk = c2s.ConvBlock([10, 64, 64], pad_value=0, norm="group").cuda().eval()
    x = torch.randn((2, 10, 6, 128), device="cuda").to(torch.bfloat16)
    c2s.ops._lib.reset_launch_count()
    with torch.no_grad():
        blk(x)
    torch.cuda.synchronize()
    # weight preparation + convolution per layer, ONE normalisation pass: the second convolution normalises its input on the
    # fly (c2s_conv_input_norm), the statistics come out of the convolutions
    assert c2s.ops._lib.launch_count() == 5


@pytest.mark.gpu
def test_input_normalisation_on_the_fly_equals_the_separate_pass():
    """conv2(relu(GroupNorm(raw1))) with the normalisation inside the second convolution's producers against the same with
    c2s_group_norm_relu in between: the arithmetic is the same (fp32 fma, ReLU, round to bf16), so are the bits."""
    g = torch.Generator(device="cuda").manual_seed(9)
    blk = c2s.ConvBlock([10, 64, 64], pad_value=0, norm="group").cuda().eval()
    with torch.no_grad():
        for m in blk.modules():
            if isinstance(m, torch.nn.GroupNorm):
                m.weight.copy_(1 + 0.3 * torch.randn(64, device="cuda", generator=g))
                m.bias.copy_(0.2 * torch.randn(64, device="cuda", generator=g))
    x = torch.randn((3, 10, 21, 128), device="cuda", generator=g).to(torch.bfloat16)
    seq = blk.conv.conv
    with torch.no_grad():
        fused = blk(x)
        raw1, st1 = c2s_conv.conv2d_reflect_forward(x, seq[0].weight, seq[0].bias)
        a1 = c2s_conv.group_norm_relu(raw1, st1, seq[1], relu=True)
        raw2, st2 = c2s_conv.conv2d_reflect_forward(a1, seq[3].weight, seq[3].bias)
        ref = c2s_conv.group_norm_relu(raw2, st2, seq[4], relu=True)
    # the GroupNorm sums are accumulated with float atomics (order-dependent last bits): compare within bf16 rounding
    assert _rel(fused.float().cpu().numpy(), ref.float().cpu().numpy()) < 8e-3
    raw2f, _ = c2s_conv.conv2d_reflect_forward(raw1, seq[3].weight, seq[3].bias, in_norm=(st1, seq[1], True))
    assert torch.equal(raw2f, raw2)  # same statistics tensor in both paths: bit-exact
