#!/usr/bin/env python
"""Webapp-style full Sentinel-2 tile through the hot path (BASELINE.json configs[4]): a 10980 x 10980 tile is padded
to 11008 x 11008 and cut into 86 x 86 = 7396 patches of 128 x 128, T = 60 frames, U-TAE
placement (LTAE on [B,60,128,16,16] + three aggregations), patches sharded contiguously over the ranks
(`crop2seg_b200.shard_bounds`), batches of 64, no collective on the data path.

    python tools/bench_tile.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_tile.py

Synthetic features: one batch of 64 patches is generated per rank and reused for every batch of the tile (the
kernels are data independent; 11 GB per batch exceeds L2, so nothing is served from cache).  One JSON line: seconds
per tile (max over ranks, CUDA events), patches/s over all ranks, fraction of the HBM roofline.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crop2seg_b200 as c2s  # noqa: E402
from crop2seg_b200 import _lib  # noqa: E402
from c2s_testlib import randomise  # noqa: E402
from bench import LEVELS, LTAE_C, LTAE_RES, N_HEAD, make_positions  # noqa: E402

T_TILE = 60
N_PATCHES = 86 * 86


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tiles", type=int, default=1)
    args = ap.parse_args()
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    lo, hi = c2s.shard_bounds(N_PATCHES, rank, world)
    lengths = np.full(B, T_TILE)  # a tile has one acquisition list: every patch sees all 60 dates
    pos_np, pad_np = make_positions(lengths, 1234)
    pos, pad = torch.from_numpy(pos_np[:, :T_TILE]).to(dev), torch.from_numpy(pad_np[:, :T_TILE]).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)

    def feat(c, r):
        x = torch.empty((B, T_TILE, c, r, r), dtype=torch.bfloat16, device=dev)
        for i in range(B):
            x[i] = torch.randn((T_TILE, c, r, r), device=dev, generator=gen).clamp_(min=0).to(torch.bfloat16)
        return x

    x4, xs = feat(LTAE_C, LTAE_RES), [feat(c, r) for c, r in LEVELS]
    enc = c2s.LTAE(in_channels=LTAE_C, n_head=N_HEAD, d_k=4, mlp=[256, 128], d_model=256)
    randomise(enc, np.random.RandomState(1234))
    enc = enc.to(dev).eval()
    agg = c2s.TemporalAggregator(mode="att_group")

    def run_tile():
        done = lo
        while done < hi:
            n = min(B, hi - done)
            with torch.no_grad():
                out, att = enc(x4[:n], batch_positions=pos[:n], pad_mask=pad[:n])
                for x in xs:
                    agg(x[:n], pad_mask=pad[:n], attn_mask=att)
            done += n

    run_tile()  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.tiles):
        run_tile()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item()) * 1e-3 / args.tiles
    e_in = LTAE_C * LTAE_RES ** 2 + sum(c * r * r for c, r in LEVELS)
    per_patch = 2 * T_TILE * e_in + 2 * e_in + 4 * N_HEAD * T_TILE * LTAE_RES ** 2
    peak_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peak_path))["hbm_gbs"]) if os.path.exists(peak_path) else 6650.0
    if rank == 0:
        print(json.dumps({
            "metric": "patches/sec, full Sentinel-2 tile (10980^2 -> 7396 patches of 128^2, T=60), LTAE+aggregator fwd",
            "value": N_PATCHES / sec, "unit": "patches/s", "n_gpus": world, "seconds_per_tile": sec, "scaling": "strong",
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE configs[4]", "patches": N_PATCHES, "batch": B, "frames": T_TILE,
                       "sharding": f"contiguous shards of {hi - lo} patches on rank 0 of {world}"},
            "roofline": {"bound": "hbm", "algorithmic_bytes": per_patch * N_PATCHES,
                         "achieved": per_patch * N_PATCHES / sec / 1e9 / world, "peak": peak, "unit": "GB/s per GPU",
                         "frac": per_patch * N_PATCHES / sec / 1e9 / world / peak},
            "gpu_launches_per_tile_rank0": _lib.launch_count() / args.tiles,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
