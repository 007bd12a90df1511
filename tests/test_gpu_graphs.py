"""The hot path is CUDA-graph capturable: no host synchronisation, no allocation outside torch's pools, tensor maps
encoded on the host at capture time.  A deployed model replays one graph per batch shape; the training step
(forward, backward, Adam) replays as one graph too (tools/bench_training.py --graph)."""
import numpy as np
import pytest
import torch

import crop2seg_b200 as c2s
from c2s_testlib import randomise, synth_inputs, to_dev

pytestmark = pytest.mark.gpu


def _setup(rng, b=3, t=13, lengths=(13, 6, 9)):
    enc = c2s.LTAE(in_channels=128, n_head=16, d_k=4, mlp=[256, 128], d_model=256)
    randomise(enc, rng)
    x4, pos, pad = synth_inputs(rng, b, t, 128, 4, 4, list(lengths))
    x1 = synth_inputs(rng, b, t, 64, 32, 32, list(lengths))[0]
    x1[np.asarray(pad)] = 0
    return enc.cuda(), to_dev(x4, dtype=torch.bfloat16), to_dev(x1, dtype=torch.bfloat16), to_dev(pos), to_dev(pad)


def test_inference_hot_path_replays_bit_exactly():
    enc, x4, x1, pos, pad = _setup(np.random.RandomState(3))
    enc.eval()
    agg = c2s.TemporalAggregator("att_group")

    def run():
        with torch.no_grad():
            out, att = enc(x4, batch_positions=pos, pad_mask=pad)
            return out, att, agg(x1, pad_mask=pad, attn_mask=att)

    eager = [t.clone() for t in run()]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = run()
    x4_new = torch.roll(x4, 1, dims=0)  # new data in the captured input buffers
    x4_saved = x4.clone()
    x4.copy_(x4_new)
    graph.replay()
    with torch.no_grad():
        again = [t.clone() for t in static]
    x4.copy_(x4_saved)
    graph.replay()
    torch.cuda.synchronize()
    for s, e in zip(static, eager):
        assert torch.equal(s, e)
    assert not torch.equal(again[0], eager[0])  # the replay really read the new input


def test_training_step_replays_in_one_graph():
    """Three replays of a captured forward + backward + Adam step equal three eager steps (dropout off so that both
    see the same masks)."""
    old = c2s.modules.ATTENTION_DROPOUT
    c2s.modules.ATTENTION_DROPOUT = 0.0
    try:
        results = []
        for use_graph in (False, True):
            enc, x4, x1, pos, pad = _setup(np.random.RandomState(5))
            enc.train()
            enc.mlp[5].p = 0.0
            enc.assume_zero_padded = True
            agg = c2s.TemporalAggregator("att_group")
            opt = torch.optim.Adam(enc.parameters(), lr=1e-3, capturable=True)
            x4.requires_grad_(True), x1.requires_grad_(True)

            def step():
                opt.zero_grad(set_to_none=True)
                x4.grad = None
                x1.grad = None
                out, att = enc(x4, batch_positions=pos, pad_mask=pad)
                loss = out.float().square().mean() + agg(x1, pad_mask=pad, attn_mask=att).float().square().mean()
                loss.backward()
                opt.step()
                return loss

            if use_graph:
                state = {k: v.clone() for k, v in enc.state_dict().items()}
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    step()
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                enc.load_state_dict(state)  # undo the warm-up step ...
                for st in opt.state.values():  # ... and Adam's: its state must exist before the capture, zeroed in place
                    st["step"].zero_(), st["exp_avg"].zero_(), st["exp_avg_sq"].zero_()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    loss = step()
                for _ in range(3):  # the capture itself executes nothing
                    graph.replay()
            else:
                for _ in range(3):
                    loss = step()
            torch.cuda.synchronize()
            results.append(({k: v.detach().float().clone() for k, v in enc.named_parameters()}, float(loss.detach())))
        (p_eager, l_eager), (p_graph, l_graph) = results
        assert np.isfinite(l_graph) and abs(l_graph - l_eager) <= 1e-3 * max(abs(l_eager), 1e-6)
        for k in p_eager:
            err = float((p_eager[k] - p_graph[k]).abs().max() / p_eager[k].abs().max().clamp_min(1e-12))
            # float atomics reorder between runs and Adam normalises the step size: an element whose gradient sits at
            # that noise level moves by +-lr per step whatever its sign does (3 steps of 1e-3 on weights of ~0.25)
            assert err < 1e-2, (k, err)
    finally:
        c2s.modules.ATTENTION_DROPOUT = old


def test_captured_graph_survives_cache_eviction_and_weight_updates():
    """A captured inference graph must not depend on the eval-mode folded-weight cache: another shape evicts the cached
    workspace, an in-place weight update makes it stale -- replays still equal an eager call with the current weights."""
    enc, x4, x1, pos, pad = _setup(np.random.RandomState(5))
    enc.eval()

    def run():
        with torch.no_grad():
            return enc(x4, batch_positions=pos, pad_mask=pad)

    run()  # warm-up fills the cache
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static = run()
    # (a) evict: an eager call with another T replaces the module's cached workspace, then allocate over the freed block
    with torch.no_grad():
        enc(x4[:, :7].contiguous(), batch_positions=pos[:, :7].contiguous(), pad_mask=pad[:, :7].contiguous())
    junk = [torch.full((1 << 20,), 7.0, device="cuda") for _ in range(8)]
    graph.replay()
    torch.cuda.synchronize()
    eager = run()
    assert torch.equal(static[0], eager[0]) and torch.equal(static[1], eager[1])
    assert all(float(j.min()) == 7.0 and float(j.max()) == 7.0 for j in junk)  # the replay wrote nowhere else
    # (b) weights change in place: the replay folds the new weights
    with torch.no_grad():
        enc.attention_head.Q.mul_(1.3)
        enc.mlp[0].weight.add_(0.02)
    graph.replay()
    torch.cuda.synchronize()
    eager2 = run()
    assert torch.equal(static[0], eager2[0]) and torch.equal(static[1], eager2[1])
    assert not torch.equal(eager2[1], eager[1])
