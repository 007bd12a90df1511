import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crop2seg_b200 import conv as cc
x = torch.randn((1024, 64, 128, 128), device='cuda').to(torch.bfloat16)
conv = torch.nn.Conv2d(64, 64, 3, padding=1, padding_mode='reflect').cuda()
for _ in range(3):
    cc.conv2d_reflect_forward(x, conv.weight, conv.bias)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    cc.conv2d_reflect_forward(x, conv.weight, conv.bias)
e.record(); torch.cuda.synchronize()
print("dbg", os.environ.get("C2S_CONV_DBG", "0"), "ms", s.elapsed_time(e) / 20)
