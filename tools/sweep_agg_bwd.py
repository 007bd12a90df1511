"""A/B of the frame-chunk split of the aggregator backward (C2S_AGG_BWD_CHUNKS) at the training placement."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from crop2seg_b200 import ops
from bench import make_lengths, make_positions, T_FRAMES
dev = torch.device('cuda', 0); B = 16
lengths = make_lengths(B, 1234); _, pad_np = make_positions(lengths, 1234)
pad = torch.from_numpy(pad_np).to(dev)
att = torch.softmax(torch.randn((16, B, T_FRAMES, 16, 16), device=dev), dim=2)
for r in (32, 64, 128):
    x = torch.randn((B, T_FRAMES, 64, r, r), device=dev).clamp_(min=0).to(torch.bfloat16); x[pad] = 0
    go = torch.randn((B, 64, r, r), device=dev).to(torch.bfloat16)
    nbytes = 2 * int(lengths.sum()) * 64 * r * r + 2 * B * T_FRAMES * 64 * r * r + 2 * B * 64 * r * r
    for rep in range(2):
        for ch in ("pipe2", "pipe3", "pipe4", "pipe6"):
            os.environ.pop('C2S_AGG_NO_PIPE', None); os.environ.pop('C2S_AGG_BWD_STAGES', None)
            if ch == "reg": os.environ['C2S_AGG_NO_PIPE'] = '1'
            else: os.environ['C2S_AGG_BWD_STAGES'] = ch[4:]
            for _ in range(3): ops.temporal_aggregate_backward(x, pad, att, go, 'att_group')
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10): ops.temporal_aggregate_backward(x, pad, att, go, 'att_group')
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 10
            print(f"res {r:3d} variant {ch:>5s}: {ms*1000:7.1f} us {nbytes/ms/1e6:6.0f} GB/s", flush=True)
