import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import crop2seg_b200 as c2s
from crop2seg_b200 import _lib
from oracle import ltae_forward
from c2s_testlib import *
from golden_util import rel_err
C, co, (b, t, h, w), lengths = 64, 128, (2, 61, 8, 8), [61, 33]
kw = dict(in_channels=C, n_head=16, d_k=4, d_model=256, mlp=[256, co])
for trial in range(6):
    rng = np.random.RandomState(3)
    m = c2s.LTAE(**kw); randomise(m, rng); m = m.cuda().eval()
    for zp in (False, True):
        m.assume_zero_padded = zp
        x, pos, pad = synth_inputs(rng, b, t, C, h, w, lengths)
        ref_o, ref_a = ltae_forward(oracle_config("ltae", kw), oracle_params(m), bf16_round(x), pos, pad)
        for rep in range(3):
            with torch.no_grad():
                o, a_ = m(to_dev(x, dtype=torch.bfloat16), batch_positions=to_dev(pos), pad_mask=to_dev(pad))
            torch.cuda.synchronize()
            an = a_.cpu().numpy()
            err = np.abs(an - ref_a).reshape(16, b, t, h * w // 8, 8).max(axis=(0, 2, 4)) / np.abs(ref_a).max()
            print(trial, zp, rep, _lib.last_ltae_kernel(), "attn %.2e" % rel_err(an, ref_a), "out %.2e" % rel_err(o.float().cpu().numpy(), ref_o),
                  "bad tiles", np.argwhere(err > 1e-2).tolist(), flush=True)
