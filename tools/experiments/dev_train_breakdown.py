"""Where does the training step of tools/bench_training.py spend its time?  CUDA events around the sections plus a
torch profiler table of the kernels of one step."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import crop2seg_b200 as c2s
from c2s_testlib import randomise
from bench import LEVELS, LTAE_C, LTAE_RES, N_HEAD, T_FRAMES, make_lengths, make_positions
dev = torch.device('cuda', 0); B = 16
lengths = make_lengths(B, 1234); pos_np, pad_np = make_positions(lengths, 1234)
pos, pad = torch.from_numpy(pos_np).to(dev), torch.from_numpy(pad_np).to(dev)
def feat(c, r):
    x = torch.randn((B, T_FRAMES, c, r, r), device=dev).clamp_(min=0); x[pad] = 0
    return x.to(torch.bfloat16).requires_grad_(True)
x4, xs = feat(LTAE_C, LTAE_RES), [feat(c, r) for c, r in LEVELS]
enc = c2s.LTAE(in_channels=LTAE_C, n_head=N_HEAD, d_k=4, mlp=[256, 128], d_model=256); randomise(enc, np.random.RandomState(1)); enc = enc.to(dev).train()
enc.assume_zero_padded = True
agg = c2s.TemporalAggregator('att_group'); opt = torch.optim.Adam(enc.parameters(), lr=1e-3)
projs = [torch.randn((B, 128, 16, 16), device=dev).bfloat16()] + [torch.randn((B, c, r, r), device=dev).bfloat16() for c, r in LEVELS]
def step(ev=None):
    def mark(name):
        if ev is not None:
            e = torch.cuda.Event(enable_timing=True); e.record(); ev.append((name, e))
    mark('start'); opt.zero_grad(set_to_none=True)
    for x in [x4] + xs: x.grad = None
    out, att = enc(x4, batch_positions=pos, pad_mask=pad); mark('ltae fwd')
    loss = (out * projs[0]).float().mean()
    for x, pr in zip(xs, projs[1:]): loss = loss + (agg(x, pad_mask=pad, attn_mask=att) * pr).float().mean()
    mark('agg fwd + loss'); loss.backward(); mark('backward'); opt.step(); mark('adam')
for _ in range(5): step()
torch.cuda.synchronize()
tot = {}
for _ in range(20):
    ev = []; step(ev); torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(ev, ev[1:]): tot[n1] = tot.get(n1, 0) + e0.elapsed_time(e1) / 20
print({k: round(v, 3) for k, v in tot.items()}, 'sum', round(sum(tot.values()), 3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))
