import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import crop2seg_b200 as c2s
from crop2seg_b200 import _lib
from oracle import ltae_forward
from c2s_testlib import *
np.set_printoptions(linewidth=250, precision=3, suppress=True)
for C, co, (b, t, h, w), lengths in [(64, 64, (2, 40, 4, 8), [40, 33]), (64, 64, (5, 61, 16, 16), [61, 27, 44, 61, 30]), (128, 128, (20, 61, 8, 8), [61] * 20)]:
    kw = dict(in_channels=C, n_head=16, d_k=4, d_model=256, mlp=[256, co])
    rng = np.random.RandomState(3)
    m = c2s.LTAE(**kw); randomise(m, rng); m = m.cuda().eval()
    x, pos, pad = synth_inputs(rng, b, t, C, h, w, lengths)
    ref_o, ref_a = ltae_forward(oracle_config("ltae", kw), oracle_params(m), bf16_round(x), pos, pad)
    with torch.no_grad():
        o, a_ = m(to_dev(x, dtype=torch.bfloat16), batch_positions=to_dev(pos), pad_mask=to_dev(pad))
    torch.cuda.synchronize()
    err = np.abs(o.float().cpu().numpy() - ref_o) / np.abs(ref_o).max()      # [b, co, h, w]
    e_tile = err.reshape(b, co, h * w // 8, 8).max(axis=(1, 3))            # [b, tiles]
    print(C, co, (b, t, h, w), _lib.last_kernel(), "max", err.max())
    print("per tile (rows = sample):\n", e_tile)
    bad = np.argwhere(err > 1e-2)
    print("bad elements", len(bad), "of", err.size, bad[:12].tolist())
    e_ch = err.max(axis=(0, 2, 3)); print("per channel:", e_ch)
