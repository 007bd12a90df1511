"""World-size-2 gloo test of the multi-GPU inference plumbing: patches are sharded
across ranks with no data-path collective and the gathered result equals the unsharded one bit for bit.
The per-rank compute is the torch-CPU oracle standing in for the CUDA kernels (there is no GPU here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import crop2seg_b200 as c2s
from oracle.torch_port import temporal_aggregator_torch
from c2s_testlib import random_attention, synth_inputs


def _inputs():
    rng = np.random.RandomState(11)
    lengths = [5, 3, 4, 5, 2]
    x, _, pad = synth_inputs(rng, 5, 5, 8, 8, 8, lengths)
    attn = random_attention(rng, 4, pad, 4, 4)
    return torch.from_numpy(x), torch.from_numpy(pad), torch.from_numpy(attn)


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, pad, attn = _inputs()
    lo, hi = c2s.shard_bounds(x.shape[0], rank, world)
    local = temporal_aggregator_torch(x[lo:hi], pad[lo:hi], attn[:, lo:hi], "att_group")
    full = c2s.gather_shards(local, world)
    if rank == 0:
        torch.save(full, out_path)
    dist.destroy_process_group()


def test_sharded_equals_unsharded(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, out_path), nprocs=2, join=True)
    x, pad, attn = _inputs()
    ref = temporal_aggregator_torch(x, pad, attn, "att_group")
    assert torch.equal(torch.load(out_path), ref)


def _bucket_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    bucket = c2s.GradientBucket(net.parameters())
    x = torch.arange(30, dtype=torch.float32).view(6, 5) / 10 + rank  # every rank owns its shard of the batch
    for _ in range(2):  # the views survive a second step
        bucket.zero()
        net(x).square().mean().backward()
        bucket.all_reduce()
    if rank == 0:
        torch.save([p.grad.clone() for p in net.parameters()], out_path)
    dist.destroy_process_group()


def test_gradient_bucket_averages_over_ranks(tmp_path):
    """One all-reduce of one flat buffer == the mean of the per-shard gradients (what DDP computes)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "grads.pt")
    mp.spawn(_bucket_worker, args=(2, port, out_path), nprocs=2, join=True)
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 2))
    ref = [torch.zeros_like(p) for p in net.parameters()]
    for rank in range(2):
        net.zero_grad()
        x = torch.arange(30, dtype=torch.float32).view(6, 5) / 10 + rank
        net(x).square().mean().backward()
        for r, p in zip(ref, net.parameters()):
            r += p.grad / 2
    for got, want in zip(torch.load(out_path), ref):
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
