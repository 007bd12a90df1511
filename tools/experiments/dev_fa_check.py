"""Development check of the persistent TMA L-TAE kernel (default) against the oracle and the older mma.sync kernel
(C2S_LTAE_MMA=1)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import crop2seg_b200 as c2s
from crop2seg_b200 import _lib
from oracle import ltae4wtae_forward, ltae_forward
from c2s_testlib import *
from golden_util import rel_err
cases = [("ltae4wtae", 128, 128, (2, 61, 8, 8), [61, 27], {}), ("ltae", 128, 128, (2, 61, 8, 8), [61, 27], {}),
         ("ltae", 128, 128, (3, 17, 4, 4), [17, 0, 1], {}), ("ltae", 64, 64, (2, 40, 4, 8), [40, 33], {}),
         ("ltae", 64, 64, (5, 61, 16, 16), [61, 27, 44, 61, 30], {}),
         ("ltae", 64, 128, (2, 61, 8, 8), [61, 33], {}),
         ("ltae", 128, 96, (2, 64, 4, 4), [64, 50], {}),
         ("ltae", 128, 128, (70, 61, 8, 8), [61] + [27 + (i * 7) % 35 for i in range(69)], {}),
         ("ltae", 128, 128, (2, 33, 4, 4), [33, 30], {"positional_encoding": False}),
         ("ltae", 128, 128, (2, 40, 4, 4), [40, 28], {"use_doy": True})]
for kind, C, co, (b, t, h, w), lengths, extra in cases:
    kw = dict(in_channels=C, n_head=16, d_k=4, d_model=256, **extra)
    if kind == "ltae": kw["mlp"] = [256, co]
    rng = np.random.RandomState(3)
    m = (c2s.LTAE if kind == "ltae" else c2s.LTAE4WTAE)(**kw); randomise(m, rng); m = m.cuda().eval()
    for zp in (False, True):
        m.assume_zero_padded = zp
        x, pos, pad = synth_inputs(rng, b, t, C, h, w, lengths, doy=bool(extra.get("use_doy")))
        if extra.get("positional_encoding") is False: pos = None
        if kind == "ltae":
            ref_o, ref_a = ltae_forward(oracle_config(kind, kw), oracle_params(m), bf16_round(x), pos, pad)
        else:
            ref_o, ref_a = None, ltae4wtae_forward(oracle_config(kind, kw), oracle_params(m), bf16_round(x), pos, pad)
        res = {}
        for old in (True, False):
            os.environ.pop('C2S_LTAE_MMA', None)
            if old: os.environ['C2S_LTAE_MMA'] = '1'
            with torch.no_grad():
                r = m(to_dev(x, dtype=torch.bfloat16), batch_positions=to_dev(pos), pad_mask=to_dev(pad))
            torch.cuda.synchronize()
            o, a_ = (r if kind == "ltae" else (None, r))
            an = a_.cpu().numpy()
            res[old] = (_lib.last_kernel(), "attn %.2e" % rel_err(an, ref_a), "sum %.1e" % np.abs(an.sum(2) - 1).max(),
                        None if o is None else "out %.2e" % rel_err(o.float().cpu().numpy(), ref_o))
        print(kind, C, co, (b, t, h, w), "zp" if zp else "  ", extra, "\n    old", res[True], "\n    new", res[False], flush=True)
