#!/usr/bin/env python
"""Hot-path throughput at the W-TAE and Time-Unet placements (BASELINE.json configs[2]); one JSON line per placement.
The same records ride in bench.py's line (``placements``).

    python tools/bench_placements.py [--batch 64] [--seconds 1.0] [--only wtae,timeunet,timeunet_att]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools import bench_lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--only", default="wtae,timeunet,timeunet_att")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    for name, rec in bench_lib.placements(dev, B=args.batch, min_seconds=args.seconds, only=tuple(args.only.split(","))).items():
        print(json.dumps({"placement": name, **rec}), flush=True)


if __name__ == "__main__":
    main()
