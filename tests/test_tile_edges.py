"""Edges of the tile inference pipeline.

CPU: ``oracle/tile_oracle.py`` against what the reference's OWN code produced (tests/golden/make_tile_golden.py ran
``DatasetCreator._patchify`` -> ``S2TSCZCropDataset.__getitem__`` -> ``pad_collate`` and the post-processing statements
of ``generate_prediction``): model inputs bit for bit (SHA-256 of the whole tensor), class map bit for bit.
GPU (``-m gpu``): the kernels behind ``c2s_tile_patchify`` / ``c2s_tile_classmap`` against the same vectors and against
the oracle on other shapes (uint16 / float32 tiles, bf16 patches, sharded patch ranges, ragged tile edges).
"""
import hashlib
import json
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_tile_golden import synth_logits, synth_tile  # noqa: E402  (seeded generators only; no reference import)
from oracle.tile_oracle import CHANNELS_LIKE_PASTIS, classmap_from_logits, patch_grid, patchify_normalise  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tile_edges.npz")


def _gold():
    z = np.load(GOLD, allow_pickle=False)
    return json.loads(str(z["cfg"])), z


def _sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_oracle_patchify_equals_the_reference_dataset_path():
    cfg, z = _gold()
    c = cfg["pre"]
    tile = synth_tile(c["seed"], c["T"], c["C"], c["H"], c["W"])
    x = patchify_normalise(tile, c["channels_order"], z["pre::mean"], z["pre::std"], t_pad=c["T_pad"],
                           pad_value=c["pad_value"], grid=tuple(c["grid"]))
    assert x.shape == (12, c["T_pad"], c["C"], 128, 128) and x.dtype == np.float32
    assert np.array_equal(x[:, :, :, ::17, ::13], z["pre::sample"])
    assert np.array_equal(_sha(x), z["pre::sha256"])  # every bit of the 39 MB the reference built
    assert tuple(c["channels_order"]) == CHANNELS_LIKE_PASTIS


def test_oracle_classmap_equals_the_reference_postprocessing():
    cfg, z = _gold()
    c = cfg["post"]
    logits = synth_logits(c["seed"], c["P"], c["K"])
    cm, pr = classmap_from_logits(logits, c["H"], c["W"], grid=tuple(c["grid"]))
    assert np.array_equal(cm, z["post::classmap"])  # ties included: the first maximum
    assert np.abs(pr[:, ::37, ::41] - z["post::proba_sample"]).max() < 2e-7  # numpy exp vs ATen exp: last bit


def test_oracle_patch_ranges_compose():
    rng = np.random.RandomState(3)
    tile = rng.randint(0, 5000, size=(2, 10, 140, 300)).astype(np.int16)
    mean, std = rng.uniform(500, 2000, 10).astype(np.float32), rng.uniform(300, 900, 10).astype(np.float32)
    full = patchify_normalise(tile, CHANNELS_LIKE_PASTIS, mean, std)
    assert patch_grid(140, 300) == (2, 3) and full.shape[0] == 6
    part = patchify_normalise(tile, CHANNELS_LIKE_PASTIS, mean, std, patch_begin=2, patch_count=3)
    assert np.array_equal(part, full[2:5])
    # zero-padding is applied to the RAW values: padded pixels normalise to -mean/std (dataset_creator.py:388)
    assert np.array_equal(full[5, 0, :, 127, 127], ((np.float32(0) - mean) / std).astype(np.float32))


# ------------------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


@gpu
def test_gpu_patchify_equals_the_reference_dataset_path():
    import crop2seg_b200 as c2s
    cfg, z = _gold()
    c = cfg["pre"]
    tile = synth_tile(c["seed"], c["T"], c["C"], c["H"], c["W"])
    p = c2s.TilePatchifier(torch.from_numpy(tile).cuda(), z["pre::mean"].tolist(), z["pre::std"].tolist(),
                           channels_order=c["channels_order"], t_pad=c["T_pad"], pad_value=c["pad_value"],
                           grid=tuple(c["grid"]))
    x = p.patches().cpu().numpy()
    assert np.array_equal(_sha(x), z["pre::sha256"])
    # a shard of the patch list = the same rows
    part = p.patches(5, 4).cpu().numpy()
    assert np.array_equal(part, x[5:9])


@gpu
@pytest.mark.parametrize("raw", ["int16", "uint16", "float32"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gpu_patchify_matches_oracle(raw, dtype):
    import crop2seg_b200 as c2s
    rng = np.random.RandomState(11)
    t, c, h, w = 4, 10, 300, 200  # ragged right and bottom edges
    if raw == "float32":
        tile = (rng.standard_normal((t, c, h, w)) * 1500).astype(np.float32)
    else:
        tile = rng.randint(0, 30000 if raw == "int16" else 65535, size=(t, c, h, w)).astype(raw)
    mean, std = rng.uniform(500, 2000, c).astype(np.float32), rng.uniform(300, 900, c).astype(np.float32)
    order = rng.permutation(c).tolist()
    ref = patchify_normalise(tile, order, mean, std, t_pad=6, pad_value=-1.5)
    dev_tile = torch.from_numpy(tile.view(np.int16) if raw == "uint16" else tile).cuda()
    if raw == "uint16":
        dev_tile = dev_tile.view(torch.uint16)
    p = c2s.TilePatchifier(dev_tile, mean.tolist(), std.tolist(), channels_order=order, t_pad=6, pad_value=-1.5,
                           dtype=dtype)
    got = p.patches()
    want = torch.from_numpy(ref).to(dtype)
    assert got.dtype == dtype and torch.equal(got.cpu(), want)  # bit for bit (bf16: round-to-nearest-even of the fp32)


@gpu
def test_gpu_classmap_equals_the_reference_postprocessing():
    import crop2seg_b200 as c2s
    cfg, z = _gold()
    c = cfg["post"]
    logits = synth_logits(c["seed"], c["P"], c["K"])
    cm = c2s.ClassMap(c["H"], c["W"], c["K"], "cuda", grid=tuple(c["grid"]))
    dev = torch.from_numpy(logits).cuda()
    for begin in range(0, c["P"], 32):  # batches of patches fill disjoint parts of the map
        cm.put(dev[begin:begin + 32], begin)
    assert np.array_equal(cm.classmap.cpu().numpy(), z["post::classmap"])
    assert np.abs(cm.proba.cpu().numpy()[:, ::37, ::41] - z["post::proba_sample"]).max() < 3e-7
    ref_cm, ref_pr = classmap_from_logits(logits, c["H"], c["W"], grid=tuple(c["grid"]))
    assert np.abs(cm.proba.cpu().numpy() - ref_pr).max() < 3e-7


@gpu
def test_gpu_classmap_bf16_logits_and_no_proba():
    import crop2seg_b200 as c2s
    rng = np.random.RandomState(5)
    h, w, k = 200, 300, 20  # 2 x 3 patches, ragged edges, more than 16 classes
    logits = torch.from_numpy((rng.standard_normal((6, k, 128, 128)) * 2).astype(np.float32)).to(torch.bfloat16)
    cm = c2s.ClassMap(h, w, k, "cuda", with_proba=False)
    cm.put(logits.cuda())
    ref_cm, _ = classmap_from_logits(logits.float().numpy(), h, w)
    assert cm.proba is None and np.array_equal(cm.classmap.cpu().numpy(), ref_cm)


@gpu
def test_gpu_tile_edges_reject_bad_arguments():
    import crop2seg_b200 as c2s
    from crop2seg_b200._lib import C2SError
    tile = torch.zeros((2, 10, 64, 64), dtype=torch.int16, device="cuda")
    p = c2s.TilePatchifier(tile, [1.0] * 10, [1.0] * 10)
    with pytest.raises(C2SError):
        p.patches(1, 1)  # one patch only
    with pytest.raises(RuntimeError):
        c2s.TilePatchifier(tile.cpu(), [1.0] * 10, [1.0] * 10)
    cm = c2s.ClassMap(64, 64, 40, "cuda")
    with pytest.raises(C2SError):
        cm.put(torch.zeros((1, 40, 128, 128), device="cuda"))  # > 32 classes: unsupported, said loudly
