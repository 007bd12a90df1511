"""Experiment: the 64 -> 64 convolution at 128^2 with and without the on-the-fly input normalisation."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crop2seg_b200 import conv as cc

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

for w in (128, 64):
    x = torch.randn((1024, 64, w, w), device="cuda").to(torch.bfloat16)
    conv = torch.nn.Conv2d(64, 64, 3, padding=1, padding_mode="reflect").cuda()
    norm = torch.nn.GroupNorm(4, 64).cuda()
    stats = cc.group_stats(x, 4)
    print(w, "plain", timed(lambda: cc.conv2d_reflect_forward(x, conv.weight, conv.bias)))
    print(w, "in_norm", timed(lambda: cc.conv2d_reflect_forward(x, conv.weight, conv.bias, in_norm=(stats, norm, True))))
    y = torch.empty_like(x)
    print(w, "norm pass", timed(lambda: cc.group_norm_relu(x, stats, norm, out=y)))
