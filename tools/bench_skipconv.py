"""Aggregation + skip convolution: fused kernel vs (aggregator kernel -> torch Conv2d/BatchNorm2d/ReLU in bf16).

    python tools/bench_skipconv.py [--batch 64] [--steps 20]
One JSON line per U-TAE skip level; times are CUDA events on the launching stream, inputs larger than L2."""
import argparse, json, os, sys
import numpy as np, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import crop2seg_b200 as c2s
from crop2seg_b200 import _lib
from bench import LEVELS, T_FRAMES, make_lengths, make_positions
from c2s_testlib import random_attention

ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=64); ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0); B = args.batch
lengths = make_lengths(B, 1234); _, pad_np = make_positions(lengths, 1234)
pad = torch.from_numpy(pad_np).to(dev)
attn = torch.from_numpy(random_attention(np.random.RandomState(1), 16, pad_np, 16, 16)).to(dev)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
agg = c2s.TemporalAggregator("att_group")

def timeit(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / args.steps

for c, r in LEVELS:
    x = torch.empty((B, T_FRAMES, c, r, r), dtype=torch.bfloat16, device=dev)
    for i in range(B):
        v = torch.randn((T_FRAMES, c, r, r), device=dev).clamp_(min=0); v[pad[i]] = 0; x[i] = v.to(torch.bfloat16)
    conv = nn.Sequential(nn.Conv2d(c, c, 1), nn.BatchNorm2d(c), nn.ReLU()).to(dev).eval()
    with torch.no_grad():
        conv[1].running_mean.normal_(0, 0.3); conv[1].running_var.uniform_(0.5, 2.0)
    conv_bf16 = nn.Sequential(nn.Conv2d(c, c, 1), nn.BatchNorm2d(c), nn.ReLU()).to(dev).eval()
    conv_bf16.load_state_dict(conv.state_dict()); conv_bf16 = conv_bf16.to(torch.bfloat16).to(memory_format=torch.contiguous_format)
    with torch.no_grad():
        t_agg = timeit(lambda: agg(x, pad_mask=pad, attn_mask=attn))
        t_unfused = timeit(lambda: conv_bf16(agg(x, pad_mask=pad, attn_mask=attn)))
        t_fused = timeit(lambda: agg.forward_skip_conv(x, pad, attn, conv))
        a = agg.forward_skip_conv(x, pad, attn, conv).float(); kname = _lib.last_kernel()
        b_ = conv(agg(x, pad_mask=pad, attn_mask=attn).float())
    err = float((a - b_).abs().max() / b_.abs().max())
    nbytes = 2 * sum(lengths) * c * r * r + 2 * B * c * r * r + 4 * 16 * sum(lengths) * 256  # x (valid frames) + out + attention
    print(json.dumps({"level": f"{c}x{r}x{r}", "batch": B, "agg_only_ms": round(t_agg, 4), "agg_then_torch_skip_conv_ms": round(t_unfused, 4),
                      "fused_ms": round(t_fused, 4), "fused_gbs": round(nbytes / t_fused / 1e6, 1), "fused_frac_of_hbm_peak": round(nbytes / t_fused / 1e6 / peak, 3),
                      "rel_err_vs_unfused_fp32_conv": err, "kernel": kname}))
