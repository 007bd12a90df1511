#!/usr/bin/env python
"""Golden vectors for the edges of the tile inference pipeline, produced by the
reference's OWN code.  Runs only in the build container.

The reference modules on this path import GIS / UI packages that are not installed (rasterio, geopandas, shapely,
streamlit, torchnet ...); none of them is touched by the functions executed here, so they are replaced by empty
stand-ins in ``sys.modules`` before the import.  What runs, unmodified, from /root/reference:

  * ``DatasetCreator._patchify(data, affine, 128)`` with ``for_inference=True``  (src/helpers/dataset_creator.py:347-388)
  * ``S2TSCZCropDataset(..., for_inference=True, norm=True, channels_like_pastis=True)[i]`` on a scratch dataset folder
    written from those patches                                                    (src/datasets/s2_ts_cz_crop.py:357-474)
  * ``pad_collate(batch, pad_value=0, max_size=T_pad)``                           (src/utils.py:20-66)
  * the post-processing statements of ``generate_prediction`` (src/webapp/prediction.py:318-320, 326-333), read from
    the file and executed on seeded logits of 100 patches.

Inputs are regenerated from seeds by the tests (numpy's legacy RandomState is stable), so the fixture only holds the
seeds, the shapes, SHA-256 digests of the reference outputs and small samples of them.

    python tests/golden/make_tile_golden.py
"""
import hashlib
import json
import os
import sys
import tempfile
import textwrap
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CROP2SEG_REFERENCE", "/root/reference")


def synth_tile(seed, t, c, h, w):
    """Raw Sentinel-2 reflectances (int16, a few exact zeros = no data)."""
    rng = np.random.RandomState(seed)
    tile = rng.randint(0, 6000, size=(t, c, h, w)).astype(np.int16)
    tile[rng.uniform(size=tile.shape) < 0.01] = 0
    return tile


def synth_logits(seed, p, k):
    rng = np.random.RandomState(seed)
    logits = (rng.standard_normal((p, k, 128, 128)) * 3).astype(np.float32)
    logits[:, 3][logits[:, 7] > 6.0] = logits[:, 7][logits[:, 7] > 6.0]  # exact ties: the first maximum must win
    return logits


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


STUB_TOPLEVEL = {"rasterio", "geopandas", "shapely", "streamlit", "torchnet", "osgeo", "cv2", "sentinelsat", "fiona",
                 "pyproj", "folium", "streamlit_folium", "leafmap", "matplotlib", "seaborn", "bs4", "lxml", "tifffile",
                 "skimage", "PIL", "thop", "fvcore", "plotly", "pystac_client", "planetary_computer", "odc", "eodag"}


def _stub_missing_modules():
    """Any import below one of the (absent) packages above yields an empty stand-in module."""
    import importlib.abc
    import importlib.machinery

    class Loader(importlib.abc.Loader):
        def create_module(self, spec):
            m = mock.MagicMock(name=spec.name)
            m.__name__, m.__path__, m.__spec__, m.__loader__ = spec.name, [], spec, self
            return m

        def exec_module(self, module):
            pass

    class Finder(importlib.abc.MetaPathFinder):
        def find_spec(self, name, path=None, target=None):
            top = name.split(".")[0]
            if top not in STUB_TOPLEVEL:
                return None
            return importlib.machinery.ModuleSpec(name, Loader(), is_package=True)

    present = set()
    for top in STUB_TOPLEVEL:
        try:
            __import__(top)
            present.add(top)
        except Exception:
            pass
    STUB_TOPLEVEL.difference_update(present)
    sys.meta_path.append(Finder())


def main():
    _stub_missing_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)  # src/global_vars.py reads ../config/config.ini relative to its own file, other modules use cwd
    try:
        for _ in range(50):  # every further absent third-party package named by an ImportError joins the stand-ins
            try:
                from src.helpers.dataset_creator import DatasetCreator
                from src.datasets.s2_ts_cz_crop import S2TSCZCropDataset
                from src.utils import pad_collate
                break
            except ModuleNotFoundError as e:
                top = (e.name or "").split(".")[0]
                if not top or top == "src" or top in STUB_TOPLEVEL:
                    raise
                STUB_TOPLEVEL.add(top)
                for name in [n for n in sys.modules if n.startswith("src.")]:
                    del sys.modules[name]  # half-imported reference modules
    finally:
        os.chdir(cwd)
    print("stand-ins for absent packages:", sorted(STUB_TOPLEVEL))

    arrays, cfg = {}, {}
    # ---------------- before the model: patchify -> dataset item -> collate -------------------------------------
    seed, t, c, h, w, t_pad = 4101, 3, 10, 202, 330, 5  # + 182 = 384 x 512: 3 x 4 patches, last row / column pure padding
    tile = synth_tile(seed, t, c, h, w)
    creator = types.SimpleNamespace(for_inference=True)
    patches, coords = DatasetCreator._patchify(creator, tile, None, 128)  # np.pad by 182: the webapp's 1098 -> 1280
    assert coords is None and patches.shape[1:] == (t, c, 128, 128)
    gh, gw = (h + 182) // 128, (w + 182) // 128
    keep = list(range(gh * gw))
    assert patches.shape[0] == gh * gw
    mean = np.array(json.load(open(os.path.join(REF, "data/inference/NORM_S2_patch.json")))["train"]["mean"], dtype=np.float64)
    std = np.array(json.load(open(os.path.join(REF, "data/inference/NORM_S2_patch.json")))["train"]["std"], dtype=np.float64)
    order = [2, 1, 0, 4, 5, 6, 3, 7, 8, 9]
    norm_values = {"mean": mean[order], "std": std[order]}  # prediction.py:244-249
    with tempfile.TemporaryDirectory() as folder:
        os.makedirs(os.path.join(folder, "DATA_S2"))
        dates = {str(i): 20190101 + 10 * i for i in range(t)}
        meta = [{"ID_PATCH": i, "dates-S2": dates, "time-series_length": t, "crs": 32633} for i in range(len(keep))]
        json.dump(meta, open(os.path.join(folder, "metadata.json"), "w"))
        for i in range(len(keep)):
            with open(os.path.join(folder, "DATA_S2", f"S2_{i}"), "wb") as f:
                np.save(f, patches[i])
        ds = S2TSCZCropDataset(folder=folder, norm=True, norm_values=norm_values, reference_date="2018-09-01",
                               channels_like_pastis=True, use_doy=False, add_ndvi=False, use_abs_rel_enc=False,
                               for_inference=True, set_type="test", cache=False)
        items = [ds[i] for i in range(len(ds))]
    batch = pad_collate(items, pad_value=0, max_size=t_pad)
    x, pos = batch
    x = x.numpy()
    assert x.shape == (len(keep), t_pad, c, 128, 128) and x.dtype == np.float32
    cfg["pre"] = dict(seed=seed, T=t, C=c, H=h, W=w, T_pad=t_pad, pad_value=0.0, channels_order=order, grid=[gh, gw])
    arrays["pre::mean"] = norm_values["mean"].astype(np.float32)
    arrays["pre::std"] = norm_values["std"].astype(np.float32)
    arrays["pre::sha256"] = np.frombuffer(bytes.fromhex(digest(x)), dtype=np.uint8)
    arrays["pre::sample"] = x[:, :, :, ::17, ::13].copy()
    arrays["pre::positions"] = pos.numpy()

    # ---------------- after the model: the statements of generate_prediction ------------------------------------
    seed2, k = 4202, 15
    logits = synth_logits(seed2, 100, k)
    src = open(os.path.join(REF, "src/webapp/prediction.py")).read().splitlines()
    per_batch = textwrap.dedent("\n".join(src[317:320]))   # lines 318-320: softmax, append proba, append first maximum
    after = textwrap.dedent("\n".join(src[325:333]))       # lines 326-333: stack, rearrange 10 x 10 patches, crop to 1098
    assert "Softmax" in per_batch and "rearrange" in after and ":1098" in after, "prediction.py moved: fix the line ranges"
    from einops import rearrange
    ns = {"torch": torch, "np": np, "rearrange": rearrange, "proba": [], "t1": []}
    for i in range(100):  # batch_size = 1 (prediction.py:194)
        ns["out"] = torch.from_numpy(logits[i:i + 1])
        exec(per_batch, ns)
    exec(after, ns)
    t1, proba = ns["t1"], ns["proba"]
    assert t1.shape == (1098, 1098) and proba.shape == (k, 1098, 1098)
    cfg["post"] = dict(seed=seed2, P=100, K=k, H=1098, W=1098, grid=[10, 10])
    arrays["post::classmap"] = t1.astype(np.uint8)
    arrays["post::proba_sha256"] = np.frombuffer(bytes.fromhex(digest(proba.astype(np.float32))), dtype=np.uint8)
    arrays["post::proba_sample"] = proba[:, ::37, ::41].astype(np.float32)
    path = os.path.join(HERE, "tile_edges.npz")
    np.savez_compressed(path, cfg=json.dumps(cfg), **arrays)
    print(f"tile_edges: {os.path.getsize(path) / 1024:.1f} KiB; ties in the class map input: "
          f"{int((logits[:, 3] == logits[:, 7]).sum())}")


if __name__ == "__main__":
    main()
