"""Sharded == unsharded, bit for bit, on the REAL kernels (SURVEY.md sections 4 and 8e).

The path shards by patch with no data-path collective.  Here the U-TAE hot path (LTAE -> three aggregations) runs on
the whole batch and on the shards ``crop2seg_b200.shard_bounds`` gives 2, 3 and 4 ranks; the shards run on different
devices when the box has them (one process, round-robin over the visible GPUs), otherwise all on cuda:0.  A second
test does the same with two real ranks under NCCL (spawned processes, ``gather_shards``) when two GPUs are visible.
"""
import os
import socket

import numpy as np
import pytest
import torch

import crop2seg_b200 as c2s
from c2s_testlib import randomise, synth_inputs

pytestmark = pytest.mark.gpu

B, T = 7, 11
LENGTHS = [11, 4, 0, 7, 11, 1, 9]


def _inputs():
    rng = np.random.RandomState(404)
    x4, pos, pad = synth_inputs(rng, B, T, 128, 8, 8, LENGTHS)
    xs = [synth_inputs(rng, B, T, 64, r, r, LENGTHS)[0] for r in (16, 32, 64)]
    enc = c2s.LTAE(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256)
    randomise(enc, np.random.RandomState(405))
    return enc.eval(), x4, xs, pos, pad


def _run(enc, x4, xs, pos, pad, dev, lo, hi):
    enc = enc.to(dev)
    enc.assume_zero_padded = True
    agg = c2s.TemporalAggregator("att_group")
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) if dt is None else \
        torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev).to(dt)  # noqa: E731
    with torch.cuda.device(dev), torch.no_grad():
        out, att = enc(t(x4, torch.bfloat16), batch_positions=t(pos), pad_mask=t(pad))
        skips = [agg(t(x, torch.bfloat16), pad_mask=t(pad), attn_mask=att) for x in xs]
    return [out.cpu(), att.cpu()] + [s.cpu() for s in skips]


def test_sharded_equals_unsharded_on_the_kernels():
    enc, x4, xs, pos, pad = _inputs()
    n_dev = torch.cuda.device_count()
    full = _run(enc, x4, xs, pos, pad, torch.device("cuda", 0), 0, B)
    for world in (2, 3, 4):
        parts = []
        for rank in range(world):
            lo, hi = c2s.shard_bounds(B, rank, world)
            parts.append(_run(enc, x4, xs, pos, pad, torch.device("cuda", rank % n_dev), lo, hi))
        got = [torch.cat([p[0] for p in parts], 0), torch.cat([p[1] for p in parts], 1)] + \
              [torch.cat([p[k] for p in parts], 0) for k in (2, 3, 4)]
        for a, b in zip(got, full):
            assert a.shape == b.shape and torch.equal(a, b), world


def _nccl_worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    enc, x4, xs, pos, pad = _inputs()
    lo, hi = c2s.shard_bounds(B, rank, world)
    local = _run(enc, x4, xs, pos, pad, dev, lo, hi)
    gathered = [c2s.gather_shards(local[k].to(dev), world) for k in (0, 2, 3, 4)]  # results only, never features
    if rank == 0:
        torch.save([g.cpu() for g in gathered], out_path)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_equal_one(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "gathered.pt")
    mp.spawn(_nccl_worker, args=(2, port, out_path), nprocs=2, join=True)
    enc, x4, xs, pos, pad = _inputs()
    full = _run(enc, x4, xs, pos, pad, torch.device("cuda", 0), 0, B)
    got = torch.load(out_path)
    for a, k in zip(got, (0, 2, 3, 4)):
        assert torch.equal(a, full[k])
