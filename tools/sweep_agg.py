"""Sweep the tuning hooks of the pipelined aggregator (C2S_AGG_CONSUMERS / C2S_AGG_STAGES) at the U-TAE skip levels."""
import os, sys, itertools, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import crop2seg_b200 as c2s
from bench import make_lengths, make_positions, T_FRAMES
dev = torch.device('cuda', 0)
B = 64
lengths = make_lengths(B, 1234)
_, pad_np = make_positions(lengths, 1234)
pad = torch.from_numpy(pad_np).to(dev)
att = torch.softmax(torch.randn((16, B, T_FRAMES, 16, 16), device=dev), dim=2)
agg = c2s.TemporalAggregator('att_group')
for r in (32, 64, 128):
    x = torch.randn((B, T_FRAMES, 64, r, r), device=dev).clamp_(min=0).to(torch.bfloat16)
    x[pad] = 0
    nbytes = 2 * int(lengths.sum()) * 64 * r * r + 2 * B * 64 * r * r + 4 * 16 * int(lengths.sum()) * 256
    for cons, st in itertools.product((0, 64, 128, 256), (0, 3, 4, 6, 8, 12)):
        if cons: os.environ['C2S_AGG_CONSUMERS'] = str(cons)
        else: os.environ.pop('C2S_AGG_CONSUMERS', None)
        if st: os.environ['C2S_AGG_STAGES'] = str(st)
        else: os.environ.pop('C2S_AGG_STAGES', None)
        for _ in range(3): agg(x, pad_mask=pad, attn_mask=att)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20): agg(x, pad_mask=pad, attn_mask=att)
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f"res {r:3d} consumers {cons:3d} stages {st:2d}: {ms*1000:7.1f} us  {nbytes/ms/1e6:6.0f} GB/s  frac {nbytes/ms/1e6/6551:.3f}", flush=True)
    del x
