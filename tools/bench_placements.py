#!/usr/bin/env python
"""Hot-path throughput at the W-TAE and Time-Unet placements (BASELINE.json configs[2]) -- a companion of
bench.py (which measures configs[1], the U-TAE placement).  One JSON line per placement.

    python tools/bench_placements.py [--batch 64] [--steps 20] [--warmup 3] [--t-valid full|ragged]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crop2seg_b200 as c2s  # noqa: E402
from crop2seg_b200 import _lib  # noqa: E402
from c2s_testlib import randomise  # noqa: E402
from bench import make_lengths, make_positions, T_FRAMES  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--t-valid", default="full", choices=["full", "ragged"])
    ap.add_argument("--only", default="", help="wtae | timeunet | timeunet_att")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.batch
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    lengths = make_lengths(B, 1234) if args.t_valid == "ragged" else np.full(B, T_FRAMES)
    pos_np, pad_np = make_positions(lengths, 1234)
    pos, pad = torch.from_numpy(pos_np).to(dev), torch.from_numpy(pad_np).to(dev)
    n_valid = int(lengths.sum())
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)

    def feat(c, r):
        x = torch.empty((B, T_FRAMES, c, r, r), dtype=torch.bfloat16, device=dev)
        for i in range(B):
            v = torch.randn((T_FRAMES, c, r, r), device=dev, generator=gen).clamp_(min=0)
            v[pad[i]] = 0
            x[i] = v.to(torch.bfloat16)
        return x

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        _lib.reset_launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(args.steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / args.steps, _lib.launch_count() // args.steps

    def report(name, ms, launches, alg_bytes, note):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        print(json.dumps({"placement": name, "metric": "patches/s", "value": B / (ms * 1e-3), "ms_per_step": ms,
                          "batch": B, "dtype": "bf16", "mean_valid_frames": float(lengths.mean()),
                          "algorithmic_bytes": alg_bytes, "achieved_gbs": gbs, "hbm_peak_gbs": peak,
                          "roofline_frac": gbs / peak, "gpu_launches_per_step": launches, "note": note}), flush=True)

    agg = c2s.TemporalAggregator("att_group")
    if args.only in ("", "wtae"):
        enc = c2s.LTAE4WTAE(in_channels=128, n_head=16, d_k=4, d_model=256)
        randomise(enc, np.random.RandomState(1))
        enc = enc.to(dev).eval()
        enc.assume_zero_padded = True
        x4, x1 = feat(128, 16), feat(64, 128)

        def wtae():
            with torch.no_grad():
                att = enc(x4, batch_positions=pos, pad_mask=pad)
                return agg(x1, pad_mask=pad, attn_mask=att)
        ms, n = timed(wtae)
        alg = 2 * n_valid * (128 * 256 + 64 * 16384) + 2 * B * 64 * 16384 + 4 * 16 * T_FRAMES * 256 * B
        report("wtae", ms, n, alg, "LTAE4WTAE[B,61,128,16,16] + TemporalAggregator x8 on [B,61,64,128,128] (wtae.py:237-242)")
        del x4, x1
    if args.only in ("", "timeunet", "timeunet_att"):
        enc = c2s.LTAE(in_channels=64, n_head=16, d_k=4, mlp=[256, 64], d_model=256)
        randomise(enc, np.random.RandomState(2))
        enc = enc.to(dev).eval()
        enc.assume_zero_padded = True
        x = feat(64, 128)
        for need_att in ((False, True) if args.only == "" else ((True,) if args.only == "timeunet_att" else (False,))):
            def tu():
                with torch.no_grad():
                    return enc(x, batch_positions=pos, pad_mask=pad, return_att=need_att)
            ms, n = timed(tu)
            alg = 2 * n_valid * 64 * 16384 + 2 * B * 64 * 16384 + (4 * 16 * T_FRAMES * 16384 * B if need_att else 0)
            report("timeunet" + ("_att" if need_att else ""), ms, n, alg,
                   "LTAE(C=64, mlp=[256,64]) on [B,61,64,128,128] (timeunet.py:178-180), attention "
                   + ("returned" if need_att else "not materialised (return_att=False)"))


if __name__ == "__main__":
    main()
