"""One U-TAE-placement L-TAE forward (B=64, ragged lengths) a few times: the launch ncu captures."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crop2seg_b200 as c2s
from c2s_testlib import randomise
from tools.bench_lib import _features, _positions
dev = torch.device("cuda", 0)
B = 64
rng = np.random.RandomState(1234); lengths = rng.randint(27, 62, size=B); lengths[0] = 61
pos_np, pad_np = _positions(lengths, 1234)
pos, pad = torch.from_numpy(pos_np).to(dev), torch.from_numpy(pad_np).to(dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
x = _features(B, 61, 128, 16, pad, dev, gen)
enc = c2s.LTAE(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256)
randomise(enc, np.random.RandomState(1)); enc = enc.to(dev).eval(); enc.assume_zero_padded = True
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    for _ in range(6):
        enc(x, batch_positions=pos, pad_mask=pad)
    torch.cuda.synchronize()
    s.record()
    for _ in range(50):
        enc(x, batch_positions=pos, pad_mask=pad)
    e.record()
torch.cuda.synchronize()
print("ltae ms per call", s.elapsed_time(e) / 50)
