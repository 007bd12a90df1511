"""Numpy restatement of the reference L-TAE forward pass.  TEST INFRASTRUCTURE ONLY.

Follows the *as-written* algorithm of Many98/Crop2Seg step by step (no algebraic
collapsing), so that the fused CUDA path -- which does collapse the projections --
is checked against the arithmetic the reference really performs:

    LTAE.forward                      src/backbones/tae.py:451-504
    LTAE4WTAE.forward                 src/backbones/tae.py:589-635
    LightweightMultiHeadAttention     src/backbones/tae.py:760-807
    ScaledDotProductAttention         src/backbones/tae.py:822-847
    PositionalEncoder                 src/backbones/positional_encoding.py:7-43
    AbsolutePositionalEncoder         src/backbones/positional_encoding.py:46-73

Parameters are passed as a ``dict`` of numpy arrays keyed exactly like the
reference ``state_dict``.  All tensors are float32; the
reductions of the normalisation layers accumulate in float64 and round once,
which is at least as accurate as ATen's row-wise moments.

Parity status: pinned against outputs of the imported reference, see
``tests/golden/make_golden.py`` and ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32
MASK_FILL = F32(-1e6)  # tae.py:831 -- the literal the reference writes into padded scores


@dataclass
class LtaeConfig:
    """Constructor arguments of ``LTAE`` (tae.py:355-372) that change the arithmetic."""

    in_channels: int = 128
    n_head: int = 16
    d_k: int = 4
    mlp: List[int] = field(default_factory=lambda: [256, 128])
    d_model: Optional[int] = 256
    T: int = 1000
    positional_encoding: bool = True
    use_abs_rel_enc: bool = False
    use_doy: bool = False
    num_queries: int = 1
    add_linear: bool = False

    @property
    def width(self) -> int:  # tae.py:398-403
        return self.d_model if self.d_model is not None else self.in_channels


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def group_norm_rows(x: np.ndarray, groups: int, weight: np.ndarray, bias: np.ndarray,
                    eps: float = 1e-5) -> np.ndarray:
    """``nn.GroupNorm(groups, C)`` applied to ``x[N, C, L]`` (tae.py:432-440, call :461/:488).

    Statistics run over the (C/groups x L) block of every row -- for the input norm L is the
    *whole* temporal axis, padded frames included.  Biased variance, eps inside the sqrt.
    """
    n, c, l = x.shape
    xg = x.reshape(n, groups, (c // groups) * l).astype(np.float64)
    mean = xg.mean(axis=2, keepdims=True)
    var = xg.var(axis=2, keepdims=True)  # biased
    xh = ((xg - mean) / np.sqrt(var + eps)).reshape(n, c, l)
    out = xh * weight.astype(np.float64)[None, :, None] + bias.astype(np.float64)[None, :, None]
    return out.astype(F32)


def sinusoid_denominator(d: int, T: int = 1000, offset: int = 0) -> np.ndarray:
    """``T ** (2 * (i // 2) / d)`` for i in [offset, offset+d) (positional_encoding.py:16-18)."""
    idx = np.arange(offset, offset + d, dtype=F32)
    expo = (F32(2.0) * np.floor_divide(idx, F32(2.0))) / F32(d)
    return np.power(np.float64(T), expo.astype(np.float64)).astype(F32)


def sinusoid_positional_encoding(positions: np.ndarray, d: int, repeat: Optional[int],
                                 T: int = 1000, fc_weight: Optional[np.ndarray] = None,
                                 fc_bias: Optional[np.ndarray] = None,
                                 denom: Optional[np.ndarray] = None) -> np.ndarray:
    """``PositionalEncoder.forward`` (positional_encoding.py:25-43).

    positions[R, T] (any real dtype) -> [R, T, d*repeat]; even columns sin, odd columns cos,
    the d-wide table tiled ``repeat`` times, optionally followed by ``fc`` (add_linear).
    """
    if denom is None:
        denom = sinusoid_denominator(d, T)
    table = positions.astype(F32)[:, :, None] / denom.astype(F32)[None, None, :]
    table = table.astype(F32)
    enc = np.empty_like(table)
    enc[:, :, 0::2] = np.sin(table[:, :, 0::2])
    enc[:, :, 1::2] = np.cos(table[:, :, 1::2])
    if repeat is not None:
        enc = np.concatenate([enc] * repeat, axis=-1)
    if fc_weight is not None:
        enc = enc @ fc_weight.T.astype(F32) + fc_bias.astype(F32)
    return enc.astype(F32)


def absolute_positional_encoding(positions: np.ndarray, fc_weight: np.ndarray, fc_bias: np.ndarray,
                                 repeat: Optional[int]) -> np.ndarray:
    """``AbsolutePositionalEncoder.forward`` (positional_encoding.py:58-73).

    one_hot(day-of-year, 365) @ W^T + b  ==  column ``W[:, doy]`` + b; day 365 is out of
    range in the reference (``F.one_hot`` raises) and is rejected here too.
    """
    doy = positions.astype(np.int64)
    if doy.min() < 0 or doy.max() > 364:
        raise RuntimeError("Class values must be smaller than num_classes.")  # F.one_hot's message
    one_hot = np.zeros(doy.shape + (365,), dtype=F32)
    np.put_along_axis(one_hot, doy[..., None], F32(1.0), axis=-1)
    enc = one_hot @ fc_weight.T.astype(F32) + fc_bias.astype(F32)
    if repeat is not None:
        enc = np.concatenate([enc] * repeat, axis=-1)
    return enc.astype(F32)


def _softmax_last(s: np.ndarray) -> np.ndarray:
    m = s.max(axis=-1, keepdims=True)
    e = np.exp((s - m).astype(F32)).astype(F32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)


def lightweight_attention(v: np.ndarray, pad_rows: Optional[np.ndarray], Q: np.ndarray,
                          k_weight: np.ndarray, k_bias: np.ndarray, n_head: int, d_k: int,
                          attn_keep: Optional[np.ndarray] = None,
                          attn_drop_p: float = 0.1) -> Tuple[np.ndarray, np.ndarray]:
    """``LightweightMultiHeadAttention`` + ``ScaledDotProductAttention`` (tae.py:760-847).

    v[R, T, D] (projected + position-encoded rows), pad_rows[R, T] bool or None,
    Q[h, n, d_k] learned master queries.  Returns (output[R, n, D], attn[h, R, n, T]).
    ``attn_keep`` (same shape as attn, 0/1) injects a dropout mask: the reference drops
    the attention *before* returning it and rescales by 1/(1-p) (tae.py:836-837).
    """
    r, t, d = v.shape
    n = Q.shape[1]
    keys = (v.reshape(r * t, d) @ k_weight.T.astype(F32) + k_bias.astype(F32)).reshape(r, t, n_head, d_k)
    keys = keys.transpose(2, 0, 1, 3)  # [h, R, T, d_k]        tae.py:768-769
    scores = np.einsum("hnk,hrtk->hrnt", Q.astype(F32), keys, dtype=F32)  # tae.py:827
    scores = (scores / F32(np.power(d_k, 0.5))).astype(F32)  # tae.py:828, temperature = sqrt(d_k)
    if pad_rows is not None:
        scores = np.where(pad_rows[None, :, None, :], MASK_FILL, scores)  # tae.py:831
    attn = _softmax_last(scores)  # tae.py:836, softmax over T
    if attn_keep is not None:
        attn = (attn * attn_keep.astype(F32) / F32(1.0 - attn_drop_p)).astype(F32)  # tae.py:837
    values = v.reshape(r, t, n_head, d // n_head).transpose(2, 0, 1, 3)  # [h, R, T, D/h]   tae.py:776
    out = np.einsum("hrnt,hrtc->hrnc", attn, values, dtype=F32)  # tae.py:839
    out = out.transpose(1, 2, 0, 3).reshape(r, n, d)  # concatenate heads        tae.py:796-798
    return out.astype(F32), attn.astype(F32)


# --------------------------------------------------------------------------------------
# whole modules
# --------------------------------------------------------------------------------------
def _rows(x: np.ndarray) -> np.ndarray:
    """[B, T, C, H, W] -> [B*H*W, T, C] (tae.py:460)."""
    b, t, c, h, w = x.shape
    return np.ascontiguousarray(x.transpose(0, 3, 4, 1, 2)).reshape(b * h * w, t, c)


def _per_pixel(a: np.ndarray, h: int, w: int) -> np.ndarray:
    """[B, T] -> [B*H*W, T] (tae.py:453-457, :476-477)."""
    b, t = a.shape
    return np.broadcast_to(a[:, None, None, :], (b, h, w, t)).reshape(b * h * w, t)


def _encode_positions(cfg: LtaeConfig, params: Dict[str, np.ndarray], positions: np.ndarray,
                      h: int, w: int) -> np.ndarray:
    d = cfg.width // cfg.n_head

    def primary(pos_rows):
        if cfg.use_doy and not cfg.add_linear:  # tae.py:407-415
            return absolute_positional_encoding(pos_rows, params["positional_encoder.fc.weight"],
                                                params["positional_encoder.fc.bias"], cfg.n_head)
        return sinusoid_positional_encoding(
            pos_rows, d, cfg.n_head, cfg.T,
            params.get("positional_encoder.fc.weight") if cfg.add_linear else None,
            params.get("positional_encoder.fc.bias") if cfg.add_linear else None,
            denom=params.get("positional_encoder.denom"))

    if cfg.use_abs_rel_enc:  # tae.py:467-474
        rel = _per_pixel(positions[..., 0], h, w)
        absd = _per_pixel(positions[..., 1], h, w)
        return primary(rel) + absolute_positional_encoding(
            absd, params["positional_encoder_abs.fc.weight"],
            params["positional_encoder_abs.fc.bias"], cfg.n_head)
    return primary(_per_pixel(positions, h, w))


def _embed(cfg: LtaeConfig, params: Dict[str, np.ndarray], x: np.ndarray,
           positions: Optional[np.ndarray], pad_mask: Optional[np.ndarray]):
    b, t, c, h, w = x.shape
    pad_rows = None if pad_mask is None else _per_pixel(pad_mask.astype(bool), h, w)
    rows = _rows(x.astype(F32))  # [R, T, C]
    normed = group_norm_rows(rows.transpose(0, 2, 1), cfg.n_head, params["in_norm.weight"],
                             params["in_norm.bias"]).transpose(0, 2, 1)  # tae.py:461
    if cfg.d_model is not None:  # tae.py:463-464, Conv1d k=1 == per-row matmul
        wc = params["inconv.weight"].reshape(cfg.d_model, c).astype(F32)
        e = (normed.reshape(-1, c) @ wc.T + params["inconv.bias"].astype(F32)).reshape(-1, t, cfg.d_model)
    else:
        e = normed
    if cfg.positional_encoding:
        e = (e + _encode_positions(cfg, params, positions, h, w)).astype(F32)
    return e.astype(F32), pad_rows


def ltae4wtae_forward(cfg: LtaeConfig, params: Dict[str, np.ndarray], x: np.ndarray,
                      positions: Optional[np.ndarray] = None, pad_mask: Optional[np.ndarray] = None,
                      attn_keep: Optional[np.ndarray] = None) -> np.ndarray:
    """``LTAE4WTAE.forward`` (tae.py:589-635): attention masks only, [h, B, T, H, W]."""
    b, t, c, h, w = x.shape
    e, pad_rows = _embed(cfg, params, x, positions, pad_mask)
    _, attn = lightweight_attention(e, pad_rows, params["attention_head.Q"],
                                    params["attention_head.fc1_k.weight"],
                                    params["attention_head.fc1_k.bias"], cfg.n_head, cfg.d_k,
                                    attn_keep=attn_keep)
    return _shape_attn(attn, cfg, b, t, h, w)


def _shape_attn(attn: np.ndarray, cfg: LtaeConfig, b: int, t: int, h: int, w: int) -> np.ndarray:
    if cfg.num_queries == 1:  # tae.py:490-493
        return np.ascontiguousarray(attn.reshape(cfg.n_head, b, h, w, t).transpose(0, 1, 4, 2, 3))
    return np.ascontiguousarray(  # tae.py:495-498
        attn.reshape(cfg.n_head, b, h, w, cfg.num_queries, t).transpose(0, 1, 4, 5, 2, 3))


def ltae_forward(cfg: LtaeConfig, params: Dict[str, np.ndarray], x: np.ndarray,
                 positions: Optional[np.ndarray] = None, pad_mask: Optional[np.ndarray] = None,
                 training: bool = False, attn_keep: Optional[np.ndarray] = None,
                 mlp_keep: Optional[np.ndarray] = None, mlp_drop_p: float = 0.2,
                 bn_momentum: float = 0.1, bn_eps: float = 1e-5):
    """``LTAE.forward`` (tae.py:451-504).

    x[B, T, C, H, W], positions[B, T] (or [B, T, 2] with use_abs_rel_enc), pad_mask[B, T] bool.
    Returns (out[B, C', H, W], attn[h, B, T, H, W]) for num_queries == 1 (the reference's other
    layouts for n > 1 are reproduced too).  ``training=True`` switches BatchNorm1d to biased batch
    statistics over all B*H*W*n rows and additionally returns the updated running statistics
    (momentum 0.1, unbiased variance) as a third element; dropout masks are injected, never drawn.
    """
    b, t, c, h, w = x.shape
    n = cfg.num_queries
    e, pad_rows = _embed(cfg, params, x, positions, pad_mask)
    o, attn = lightweight_attention(e, pad_rows, params["attention_head.Q"],
                                    params["attention_head.fc1_k.weight"],
                                    params["attention_head.fc1_k.bias"], cfg.n_head, cfg.d_k,
                                    attn_keep=attn_keep)
    r = o.shape[0]
    c_out = cfg.mlp[1]
    y = o.reshape(r * n, cfg.width) @ params["mlp.0.weight"].T.astype(F32) + params["mlp.0.bias"].astype(F32)
    new_stats = None
    if training:  # BatchNorm1d over [R, C', n] (tae.py:444-446): statistics over (R, n)
        y64 = y.astype(np.float64)
        mean = y64.mean(axis=0)
        var_b = y64.var(axis=0)
        cnt = y.shape[0]
        var_u = var_b * cnt / max(cnt - 1, 1)
        new_stats = (
            ((1 - bn_momentum) * params["mlp.2.running_mean"] + bn_momentum * mean).astype(F32),
            ((1 - bn_momentum) * params["mlp.2.running_var"] + bn_momentum * var_u).astype(F32),
        )
        y = (y64 - mean) / np.sqrt(var_b + bn_eps)
    else:
        y = (y.astype(np.float64) - params["mlp.2.running_mean"]) / np.sqrt(
            params["mlp.2.running_var"].astype(np.float64) + bn_eps)
    y = (y * params["mlp.2.weight"] + params["mlp.2.bias"]).astype(F32)
    y = np.maximum(y, F32(0.0))  # tae.py:447
    if mlp_keep is not None:  # tae.py:448
        y = (y * mlp_keep.reshape(y.shape).astype(F32) / F32(1.0 - mlp_drop_p)).astype(F32)
    y = y.reshape(r, n, c_out)
    y = group_norm_rows(y.transpose(0, 2, 1), cfg.n_head, params["out_norm.weight"],
                        params["out_norm.bias"]).transpose(0, 2, 1)  # tae.py:488
    if n == 1:
        out = np.ascontiguousarray(y.reshape(b, h, w, c_out).transpose(0, 3, 1, 2))  # tae.py:494
    else:
        out = np.ascontiguousarray(y.reshape(b, h, w, n, c_out).transpose(0, 3, 4, 1, 2))  # tae.py:499
    attn = _shape_attn(attn, cfg, b, t, h, w)
    if training:
        return out, attn, new_stats
    return out, attn
