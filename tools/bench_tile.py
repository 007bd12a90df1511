#!/usr/bin/env python
"""Webapp-style full Sentinel-2 tile through the hot path (BASELINE.json configs[4]); see tools/bench_lib.py::tile.
The same record rides in bench.py's line (``tile``: Time-Unet placement, ``tile_utae``).

    python tools/bench_tile.py [--placement timeunet|utae] [--no-edges]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_tile.py
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import bench_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--tiles", type=int, default=1)
    ap.add_argument("--placement", default="timeunet", choices=["timeunet", "utae"])
    ap.add_argument("--no-edges", action="store_true")
    args = ap.parse_args()
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = bench_lib.tile(dev, args.placement, B=args.batch, tiles=args.tiles, with_edges=not args.no_edges)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
