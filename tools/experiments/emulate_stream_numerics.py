"""Numerics emulation (CPU, numpy) of the streaming L-TAE kernel's arithmetic against the oracle, before the kernel
existed: features normalised to fp16 (x_hat), score weights U in fp16 (one or two terms), probabilities in fp16 (one or
two terms), fp32 accumulation everywhere, everything after the value sums in fp32.  Prints max relative errors of the
attention and of the output for the test weights (c2s_testlib.randomise) and for weight_init-scale weights."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import crop2seg_b200 as c2s  # noqa: E402
from oracle import LtaeConfig, ltae_forward  # noqa: E402
from oracle.ltae_oracle import _encode_positions, group_norm_rows  # noqa: E402
from c2s_testlib import bf16_round, oracle_params, randomise, synth_inputs  # noqa: E402

F32 = np.float32


def f16(a):
    return a.astype(np.float16).astype(F32)


def split16(a):
    hi = f16(a)
    return hi, f16(a - hi)


def emulate(cfg, P, x, pos, pad, u_terms=1, p_terms=1, xhat_fmt="f16"):
    b, t, c, hh, ww = x.shape
    h, dk, D = cfg.n_head, cfg.d_k, cfg.d_model
    n = b * hh * ww
    rows = np.ascontiguousarray(x.transpose(0, 3, 4, 1, 2)).reshape(n, t, c).astype(F32)
    g = rows.reshape(n, t, h, c // h).astype(np.float64)
    mean = g.mean(axis=(1, 3))
    var = g.var(axis=(1, 3))
    rstd = (1.0 / np.sqrt(var + 1e-5)).astype(F32)
    mur = (mean * rstd).astype(F32)
    xh = rows.reshape(n, t, h, c // h) * rstd[:, None, :, None] - mur[:, None, :, None]
    xh = xh.reshape(n, t, c).astype(F32)
    xh = f16(xh) if xhat_fmt == "f16" else bf16_round(xh)
    wc = P["inconv.weight"].reshape(D, c).astype(np.float64)
    q = P["attention_head.Q"].reshape(h, dk).astype(np.float64)
    wk = P["attention_head.fc1_k.weight"].reshape(h, dk, D).astype(np.float64)
    bk = P["attention_head.fc1_k.bias"].reshape(h, dk).astype(np.float64)
    gamma, beta = P["in_norm.weight"].astype(np.float64), P["in_norm.bias"].astype(np.float64)
    qk = np.einsum("hj,hjd->hd", q, wk) / np.sqrt(dk)
    U = (qk @ wc) * gamma[None, :]                                   # [h, C]
    wb = P["inconv.bias"].astype(np.float64) + wc @ beta
    pe = _encode_positions(cfg, P, pos, 1, 1).reshape(b, t, D).astype(np.float64)
    cpos = (pe @ qk.T) + (qk @ wb)[None, None, :] + ((q * bk).sum(1) / np.sqrt(dk))[None, None, :]  # [B,T,h]
    LOG2E = 1.4426950408889634
    Ul = (U * LOG2E).astype(F32)
    uh, ul = split16(Ul)
    s = np.einsum("ntc,hc->nht", xh, uh, dtype=F32)
    if u_terms == 2:
        s = s + np.einsum("ntc,hc->nht", xh, ul, dtype=F32)
    s = s + (np.repeat(cpos, hh * ww, axis=0).transpose(0, 2, 1) * LOG2E).astype(F32)
    padr = np.repeat(pad, hh * ww, axis=0)
    s = np.where(padr[:, None, :], F32(-1e6 * LOG2E), s)
    e = np.exp2(s - s.max(axis=2, keepdims=True)).astype(F32)
    a = (e / e.sum(axis=2, keepdims=True)).astype(F32)              # [n, h, t]
    ph, pl = split16(a)
    z = np.einsum("nht,ntc->nhc", ph, xh, dtype=F32)
    if p_terms == 2:
        z = z + np.einsum("nht,ntc->nhc", pl, xh, dtype=F32)
    sa = a.sum(axis=2)
    zn = z.astype(np.float64) * gamma[None, None, :] + beta[None, None, :] * sa[:, :, None]
    dh = D // h
    o = np.einsum("hic,nhc->nhi", wc.reshape(h, dh, c), zn) + sa[:, :, None] * P["inconv.bias"].reshape(h, dh)[None]
    o = o + np.einsum("nht,nthi->nhi", a.astype(np.float64), np.repeat(pe, hh * ww, axis=0).reshape(n, t, h, dh))
    o = o.reshape(n, D)
    y = o @ P["mlp.0.weight"].T.astype(np.float64) + P["mlp.0.bias"]
    y = (y - P["mlp.2.running_mean"]) / np.sqrt(P["mlp.2.running_var"].astype(np.float64) + 1e-5) * P["mlp.2.weight"] + P["mlp.2.bias"]
    y = np.maximum(y, 0).astype(F32)
    y = group_norm_rows(y[:, :, None], h, P["out_norm.weight"], P["out_norm.bias"])[:, :, 0]
    out = y.reshape(b, hh, ww, -1).transpose(0, 3, 1, 2)
    attn = a.reshape(b, hh, ww, h, t).transpose(3, 0, 4, 1, 2)
    return out, attn


def rel(a, r):
    return float(np.abs(a - r).max() / np.abs(r).max())


def weight_init_scale(m, rng):
    """weight_init.py: N(0,1) Conv1d weights, Xavier Linear, N(0,1)... BN scale -- larger scores than `randomise`."""
    with torch.no_grad():
        m.inconv.weight.copy_(torch.from_numpy(rng.standard_normal(tuple(m.inconv.weight.shape)).astype(F32)))
        m.inconv.bias.copy_(torch.from_numpy(rng.standard_normal(tuple(m.inconv.bias.shape)).astype(F32)))


def main():
    for name, C, co in (("utae C=128", 128, 128), ("timeunet C=64", 64, 64)):
        for wname in ("randomise", "weight_init"):
            rng = np.random.RandomState(5)
            kw = dict(in_channels=C, n_head=16, d_k=4, mlp=[256, co], d_model=256)
            m = c2s.LTAE(**kw)
            randomise(m, rng)
            if wname == "weight_init":
                weight_init_scale(m, rng)
            P = oracle_params(m)
            x, pos, pad = synth_inputs(rng, 3, 61, C, 6, 6, [61, 27, 44])
            x = bf16_round(x)
            cfg = LtaeConfig(**kw)
            ref_out, ref_attn = ltae_forward(cfg, P, x, pos, pad)
            smax = None
            for (ut, pt, fmt) in ((1, 1, "f16"), (2, 1, "f16"), (1, 2, "f16"), (2, 2, "f16"), (1, 1, "bf16")):
                out, attn = emulate(cfg, P, x, pos, pad, ut, pt, fmt)
                print(f"{name:14s} {wname:11s} x_hat {fmt:4s} U terms {ut} p terms {pt}: attn {rel(attn, ref_attn):.2e}  "
                      f"out {rel(bf16_round(out.astype(F32)), ref_out):.2e} (fp32 out {rel(out, ref_out):.2e})")


if __name__ == "__main__":
    main()
