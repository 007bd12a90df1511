"""Experiment: bandwidth of the frame packing kernels (gather / scatter) on [B*T, 64, 128, 128] bf16 with ragged lengths."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crop2seg_b200 import staging

def timed(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

B, T = 16, 61
rng = np.random.RandomState(0)
lengths = rng.randint(27, 62, size=B)
pad = torch.zeros((B, T), dtype=torch.bool)
for i, L in enumerate(lengths): pad[i, L:] = True
pad = pad.cuda()
for c in (10, 64):
    flat = torch.randn((B * T, c, 128, 128), device="cuda").to(torch.bfloat16)
    slot, n_valid = staging.frame_slots(pad)
    nv = int(lengths.sum())
    packed = staging.gather_frames(flat, slot, nv)
    fb = c * 128 * 128 * 2
    ms = timed(lambda: staging.gather_frames(flat, slot, nv))
    print(f"gather  C={c}: {ms*1e3:8.1f} us  {2*nv*fb/ms/1e6:7.0f} GB/s")
    ms = timed(lambda: staging.scatter_frames(packed, slot, 0.0))
    print(f"scatter C={c}: {ms*1e3:8.1f} us  {(nv*fb + B*T*fb)/ms/1e6:7.0f} GB/s")
    ms = timed(lambda: staging.frame_slots(pad))
    print(f"index: {ms*1e3:8.1f} us")
    del flat, packed
