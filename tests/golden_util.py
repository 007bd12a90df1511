"""Helpers to read the committed golden fixtures (tests/golden/*.npz)."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fixture_names(prefixes):
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith(tuple(prefixes))]


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    cfg = json.loads(str(z["cfg"]))
    params = {k[len("param::"):]: z[k] for k in z.files if k.startswith("param::")}
    outs = {k[len("out::"):]: z[k] for k in z.files if k.startswith("out::")}
    inputs = {k: z[k] for k in z.files if "::" not in k and k != "cfg"}
    return cfg, inputs, params, outs


def load_grads(name):
    """``grad::*`` arrays of a training fixture (tests/golden/make_train_golden.py): reference autograd results."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k[len("grad::"):]: z[k] for k in z.files if k.startswith("grad::")}


def rel_err(a, b):
    """max|a-b| / max|b| -- the tolerance definition of SURVEY.md section 8d."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
