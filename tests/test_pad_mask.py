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
