#!/usr/bin/env python
"""Training throughput of the hot path at the U-TAE placement (BASELINE.json configs[3]): forward + backward of
``LTAE`` (training mode: BatchNorm batch statistics, both dropouts) on x[B,61,128,16,16] and of three
``TemporalAggregator('att_group')`` on x[B,61,64,{32,64,128}^2], B = 16 patches per GPU, irregular T 27..61.

    python tools/bench_training.py [--batch 16] [--steps 20] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/bench_training.py ...

Every rank owns its shard of patches (weak scaling); the encoder is wrapped in DistributedDataParallel, so the only
collective is NCCL's gradient all-reduce of its 83 k parameters (the aggregators have none).  The gradients of the
four outputs are fixed random tensors (what the conv decoder's backward would hand over; decoder and loss are outside the path).  One JSON
line: patches/s over all ranks, ms per step (max over ranks, CUDA events), the algorithmic bytes of a step
(forward + x re-read + grad_x written + output gradients) against the measured HBM peak, kernel launches per step.
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
from bench import LEVELS, LTAE_C, LTAE_RES, N_HEAD, T_FRAMES, algorithmic_bytes, make_lengths, make_positions  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--graph", action="store_true",
                    help="capture the whole step (forward, backward, Adam) in one CUDA graph and replay it (1 GPU)")
    args = ap.parse_args()
    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    seed = 1234 + rank
    lengths = make_lengths(B, seed)
    pos_np, pad_np = make_positions(lengths, seed)
    pos, pad = torch.from_numpy(pos_np).to(dev), torch.from_numpy(pad_np).to(dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    def feat(c, r):
        x = torch.empty((B, T_FRAMES, c, r, r), dtype=torch.bfloat16, device=dev)
        for i in range(B):
            v = torch.randn((T_FRAMES, c, r, r), device=dev, generator=gen).clamp_(min=0)
            v[pad[i]] = 0
            x[i] = v.to(torch.bfloat16)
        return x.requires_grad_(True)  # the conv encoder behind them needs grad_x

    x4 = feat(LTAE_C, LTAE_RES)
    xs = [feat(c, r) for c, r in LEVELS]
    enc = c2s.LTAE(in_channels=LTAE_C, n_head=N_HEAD, d_k=4, mlp=[256, 128], d_model=256)
    randomise(enc, np.random.RandomState(1234))
    enc = enc.to(dev).train()
    enc.assume_zero_padded = True
    model = torch.nn.parallel.DistributedDataParallel(enc, device_ids=[local_rank]) if world > 1 else enc
    agg = c2s.TemporalAggregator(mode="att_group")
    opt = torch.optim.Adam(enc.parameters(), lr=1e-3, capturable=args.graph, fused=not args.graph)  # train.py: Adam, lr 1e-3
    projs = [torch.randn((B, 128, LTAE_RES, LTAE_RES), device=dev, generator=gen).to(torch.bfloat16)] + \
            [torch.randn((B, c, r, r), device=dev, generator=gen).to(torch.bfloat16) for c, r in LEVELS]

    def step():
        opt.zero_grad(set_to_none=True)
        for x in [x4] + xs:
            x.grad = None
        out, att = model(x4, batch_positions=pos, pad_mask=pad)
        outs = [out] + [agg(x, pad_mask=pad, attn_mask=att) for x in xs]
        # the decoder and the loss are outside the path: its backward hands these four gradients over (fixed tensors here)
        torch.autograd.backward(outs, projs)
        opt.step()
        return outs[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = None
    if args.graph:
        if world > 1:
            raise SystemExit("--graph is a single-GPU option (DDP's bucket hooks are not captured here)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # torch's capture recipe: warm up on a side stream first
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _lib.reset_launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_loss = step()
        launches_per_step = _lib.launch_count()
        eager_step = step

        def step():  # noqa: F811
            graph.replay()
            return static_loss

    for _ in range(args.warmup):
        step()
    barrier()
    _lib.reset_launch_count()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.steps):
        loss = step()
    e.record()
    barrier()
    launches = _lib.launch_count() if launches_per_step is None else launches_per_step * args.steps
    t = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / args.steps
    peak_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peak_path))["hbm_gbs"]) if os.path.exists(peak_path) else 6650.0
    e_in = LTAE_C * LTAE_RES ** 2 + sum(c * r * r for c, r in LEVELS)
    fwd = algorithmic_bytes(lengths, 2)
    n_valid = int(lengths.sum())
    attn_bytes = 4 * N_HEAD * T_FRAMES * LTAE_RES ** 2 * B
    bwd = 2 * n_valid * e_in + 2 * B * T_FRAMES * e_in + 2 * B * e_in + 2 * attn_bytes  # x again, grad_x, grad_out, attn + grad_attn
    if rank == 0:
        print(json.dumps({
            "metric": "patches/sec (T=61,C=10,128^2) LTAE+aggregator fwd+bwd (training)", "value": world * B / (ms * 1e-3),
            "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE configs[3] hot path: LTAE(train) + 3x TemporalAggregator forward+backward, "
                                   "Adam step, DDP gradient all-reduce" + (", whole step replayed as one CUDA graph" if args.graph else ""),
                       "batch_per_gpu": B,
                       "mean_valid_frames": float(np.mean(lengths))},
            "roofline": {"bound": "hbm", "algorithmic_bytes": fwd + bwd, "achieved": (fwd + bwd) / (ms * 1e-3) / 1e9,
                         "peak": peak, "unit": "GB/s", "frac": (fwd + bwd) / (ms * 1e-3) / 1e9 / peak},
            "gpu_launches_per_step": launches / args.steps, "out_mean": float(loss.detach().float().mean()),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
