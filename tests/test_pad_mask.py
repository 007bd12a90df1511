"""pad_mask = (input == pad_value).all(-1).all(-1).all(-1) (utae.py:201-203): oracle vs the mask the
reference model derived from its raw input (model-level fixture), CUDA early-exit scan vs oracle."""
import numpy as np
import pytest
import torch

from golden_util import load
from oracle import pad_mask_from_input


def test_oracle_matches_the_reference_model():
    _, inp, _, _ = load("model_utae")
    assert np.array_equal(pad_mask_from_input(inp["raw_input"], 0.0), inp["pad_mask"])
    assert inp["pad_mask"].any() and not inp["pad_mask"].all()


@pytest.mark.gpu
def test_cuda_matches_the_reference_model():
    import crop2seg_b200 as c2s
    _, inp, _, _ = load("model_utae")
    x = torch.from_numpy(inp["raw_input"]).cuda()
    m = c2s.pad_mask_from_input(x)
    assert m.dtype == torch.bool and tuple(m.shape) == inp["pad_mask"].shape
    assert np.array_equal(m.cpu().numpy(), inp["pad_mask"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,pad_value", [((3, 7, 10, 32, 32), 0.0), ((2, 5, 3, 5, 7), 0.0), ((2, 4, 10, 128, 128), -1.0),
                                             ((1, 3, 1, 1, 1), 0.0)])
def test_cuda_matches_oracle(shape, pad_value, dtype):
    """Padded frames, a frame whose only non-pad value is its very last element, a NaN frame, odd sizes (scalar path)."""
    import crop2seg_b200 as c2s
    rng = np.random.RandomState(sum(shape))
    x = rng.standard_normal(shape).astype(np.float32)
    b, t = shape[:2]
    x[0, t - 1] = pad_value                      # padded
    x[b - 1, 0] = pad_value                      # padded ...
    x[b - 1, 0].reshape(-1)[-1] = 1.5            # ... except for the last element: NOT padded
    if t > 2:
        x[0, 1] = pad_value
        x[0, 1].reshape(-1)[x[0, 1].size // 2] = np.nan  # NaN != pad_value
    xt = torch.from_numpy(x).to(dtype).cuda()
    ref = pad_mask_from_input(xt.float().cpu().numpy(), pad_value)
    got = c2s.pad_mask_from_input(xt, pad_value).cpu().numpy()
    assert np.array_equal(got, ref)
    assert ref[0, t - 1] and not ref[b - 1, 0]
    # a view with an odd element offset takes the scalar path
    if np.prod(shape[2:]) > 1:
        flat = torch.zeros(xt.numel() + 1, dtype=dtype, device="cuda")
        flat[1:] = xt.reshape(-1)
        shifted = flat[1:].view(shape)
        assert np.array_equal(c2s.pad_mask_from_input(shifted, pad_value).cpu().numpy(), ref)


@pytest.mark.gpu
def test_smart_forward_matches_the_reference_block():
    """temp_shared_block.py:18-47 on a padded batch (fixture from the reference class), with our scan kernel, with a
    caller-supplied mask, and on a batch without padding."""
    import crop2seg_b200 as c2s
    cfg, inp, params, outs = load("smart_forward")
    conv = torch.nn.Conv2d(3, 5, kernel_size=4, stride=2, padding=1)
    conv.load_state_dict({k[len("conv."):]: torch.from_numpy(v) for k, v in params.items()})
    conv = conv.cuda().eval()
    fwd = lambda z: torch.relu(conv(z))  # noqa: E731
    x = torch.from_numpy(inp["x"]).cuda()
    with torch.no_grad():
        out = c2s.smart_forward(fwd, x, pad_value=cfg["pad_value"])
        out2 = c2s.smart_forward(fwd, x, pad_value=cfg["pad_value"], pad_mask=c2s.pad_mask_from_input(x))
        full = c2s.smart_forward(fwd, x + 1.0, pad_value=cfg["pad_value"])
    assert tuple(out.shape) == outs["out"].shape
    assert np.abs(out.cpu().numpy() - outs["out"]).max() < 1e-5
    assert torch.equal(out, out2)
    assert tuple(full.shape) == outs["out"].shape
    pad = pad_mask_from_input(inp["x"], 0.0)
    assert np.all(out.cpu().numpy()[pad] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("frame_shape", [(3, 8, 8), (5, 3, 3), (1, 1, 1)])
def test_frame_packing_kernels_match_boolean_indexing(dtype, frame_shape):
    """c2s_frame_index / gather / scatter == `out[~pad_mask]` and `temp[~pad_mask] = ...` (temp_shared_block.py:30-40),
    bit for bit, with 16-byte and odd-sized frames, more frames than one scan round, all-valid and all-padded masks."""
    import crop2seg_b200 as c2s
    rng = np.random.RandomState(17)
    n = 2500
    pad = rng.uniform(size=n) < 0.3
    pad[:40] = True
    pad[-1] = False
    x = torch.from_numpy(rng.standard_normal((n,) + frame_shape).astype(np.float32)).to(dtype).cuda()
    mask = torch.from_numpy(pad).cuda()
    slot, count = c2s.frame_slots(mask)
    n_valid = int((~pad).sum())
    assert int(count.item()) == n_valid
    ref_slot = np.where(pad, -1, np.cumsum(~pad) - 1).astype(np.int32)
    assert np.array_equal(slot.cpu().numpy(), ref_slot)
    packed = c2s.gather_frames(x, slot, n_valid)
    assert torch.equal(packed, x[~mask])
    out = c2s.scatter_frames(packed * 2, slot, -3.5)
    want = torch.full_like(x, -3.5)
    want[~mask] = x[~mask] * 2
    assert torch.equal(out, want)
    for m in (np.zeros(70, bool), np.ones(70, bool)):
        s2, c2 = c2s.frame_slots(torch.from_numpy(m).cuda())
        assert int(c2.item()) == int((~m).sum())
        assert np.array_equal(s2.cpu().numpy(), np.where(m, -1, np.arange(70)).astype(np.int32))


@pytest.mark.gpu
def test_smart_forward_without_host_sync_and_with_gradients():
    """With the lengths known on the host (pad_collate) smart_forward never reads the device; its gradients equal the
    ones of boolean indexing."""
    import crop2seg_b200 as c2s
    cfg, inp, params, outs = load("smart_forward")
    conv = torch.nn.Conv2d(3, 5, kernel_size=4, stride=2, padding=1)
    conv.load_state_dict({k[len("conv."):]: torch.from_numpy(v) for k, v in params.items()})
    conv = conv.cuda()
    fwd = lambda z: torch.relu(conv(z))  # noqa: E731
    x = torch.from_numpy(inp["x"]).cuda()
    pad = pad_mask_from_input(inp["x"], 0.0)
    lengths = (~pad).sum(axis=1).tolist()
    out = c2s.smart_forward(fwd, x, pad_value=cfg["pad_value"], pad_mask=torch.from_numpy(pad).cuda(), lengths=lengths)
    assert np.abs(out.detach().cpu().numpy() - outs["out"]).max() < 1e-5
    w = torch.randn_like(out)
    (out * w).sum().backward()
    g_kernel = conv.weight.grad.clone()
    conv.weight.grad = None
    b, t = x.shape[:2]
    flat = x.view(b * t, *x.shape[2:])
    m = torch.from_numpy(pad.reshape(-1)).cuda()
    temp = torch.ones((b * t,) + tuple(out.shape[2:]), device="cuda") * cfg["pad_value"]
    temp[~m] = fwd(flat[~m])  # the reference's statements (temp_shared_block.py:30-40)
    (temp.view_as(out) * w).sum().backward()
    assert torch.allclose(conv.weight.grad, g_kernel, rtol=1e-5, atol=1e-6)
