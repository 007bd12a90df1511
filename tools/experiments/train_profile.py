"""Experiment: torch.profiler table of one training step of the hot path (which torch ops launch the small kernels)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crop2seg_b200 as c2s
from c2s_testlib import randomise
from tools.bench_lib import _features, _positions, LEVELS, LTAE_C, LTAE_RES, N_HEAD, T_FRAMES
dev = torch.device("cuda", 0)
B = 16
rng = np.random.RandomState(1234); lengths = rng.randint(27, T_FRAMES + 1, size=B); lengths[0] = T_FRAMES
pos_np, pad_np = _positions(lengths, 1234)
pos, pad = torch.from_numpy(pos_np).to(dev), torch.from_numpy(pad_np).to(dev)
gen = torch.Generator(device=dev); gen.manual_seed(1)
x4 = _features(B, T_FRAMES, LTAE_C, LTAE_RES, pad, dev, gen, True)
xs = [_features(B, T_FRAMES, c, r, pad, dev, gen, True) for c, r in LEVELS]
enc = c2s.LTAE(in_channels=LTAE_C, n_head=N_HEAD, d_k=4, mlp=[256, 128], d_model=256)
randomise(enc, np.random.RandomState(1)); enc = enc.to(dev).train(); enc.assume_zero_padded = True
bucket = c2s.GradientBucket(enc.parameters())
agg = c2s.TemporalAggregator(mode="att_group")
opt = torch.optim.Adam(enc.parameters(), lr=1e-3, fused=True)
projs = [torch.randn((B, 128, LTAE_RES, LTAE_RES), device=dev, generator=gen).to(torch.bfloat16)] + \
        [torch.randn((B, c, r, r), device=dev, generator=gen).to(torch.bfloat16) for c, r in LEVELS]
def step():
    bucket.zero()
    for x in [x4] + xs: x.grad = None
    out, att = enc(x4, batch_positions=pos, pad_mask=pad)
    outs = [out] + [agg(x, pad_mask=pad, attn_mask=att) for x in xs]
    torch.autograd.backward(outs, projs)
    bucket.all_reduce()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=48, max_shapes_column_width=60))
