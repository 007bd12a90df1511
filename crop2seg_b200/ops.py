"""Tensor-level entry points over the C ABI (``include/crop2seg_b200.h``).

PyTorch is plumbing here: it owns device memory and the current stream; every arithmetic step runs
in ``libcrop2seg_b200.so``.  Inputs must live on a CUDA device -- there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_AGG_MODES = {"att_group": _lib.AGG_ATT_GROUP, "att_mean": _lib.AGG_ATT_MEAN, "mean": _lib.AGG_MEAN}
_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
_FOLDED_CACHE_MAX_BYTES = 64 << 20  # larger L-TAE workspaces are transient (their preparation is noise there)


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"crop2seg_b200: {name} is on {t.device}; the operators are CUDA (sm_100a) only and have no CPU fallback")


def _dtype_code(t: torch.Tensor, name: str) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"crop2seg_b200: {name} has dtype {t.dtype}; supported: float32, bfloat16") from None


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _mask_u8(pad_mask: Optional[torch.Tensor], b: int, t: int, device) -> Optional[torch.Tensor]:
    if pad_mask is None:
        return None
    if tuple(pad_mask.shape) != (b, t):
        raise RuntimeError(f"crop2seg_b200: pad_mask has shape {tuple(pad_mask.shape)}, expected {(b, t)}")
    m = pad_mask.to(device=device)
    if m.dtype != torch.bool:
        m = m != 0
    return m.contiguous().view(torch.uint8)


def pad_mask_from_input(x: torch.Tensor, pad_value: float = 0.0) -> torch.Tensor:
    """``(x == pad_value).all(dim=-1).all(dim=-1).all(dim=-1)`` for x[B,T,C,H,W] (utae.py:201-203, wtae.py:221-223,
    timeunet.py:170-172): bool [B,T].  One kernel; a valid frame is left at its first non-pad value, only padded
    frames are read to the end (the reference compares and reduces the whole input three times)."""
    if x.dim() != 5:
        raise RuntimeError(f"crop2seg_b200: x must be [B,T,C,H,W], got {tuple(x.shape)}")
    _require_cuda(x, "x")
    x = x.contiguous()
    b, t = x.shape[:2]
    mask = torch.empty((b, t), dtype=torch.uint8, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        status = lib.c2s_pad_mask(x.data_ptr(), _dtype_code(x, "x"), b * t, x[0, 0].numel(), float(pad_value),
                                  mask.data_ptr(), _stream(x.device))
    _lib.check(status, "c2s_pad_mask")
    return mask.view(torch.bool)


def temporal_aggregate(x: torch.Tensor, pad_mask: Optional[torch.Tensor] = None,
                       attn_mask: Optional[torch.Tensor] = None, mode: str = "mean") -> torch.Tensor:
    """``TemporalAggregator(mode).forward(x, pad_mask, attn_mask)`` (temporal_aggregator.py:14-77), differentiable
    w.r.t. x and attn_mask (both directions are CUDA kernels)."""
    needs_grad = torch.is_grad_enabled() and (x.requires_grad or (attn_mask is not None and attn_mask.requires_grad))
    if not needs_grad:
        return temporal_aggregate_forward(x, pad_mask, attn_mask, mode)
    from .autograd import AggregateFunction
    if mode not in _AGG_MODES:
        raise ValueError(f"unknown aggregation mode {mode!r}")
    attn = None
    if mode != "mean":
        if attn_mask is None:
            raise RuntimeError(f"crop2seg_b200: mode {mode!r} needs attn_mask")
        attn = attn_mask if attn_mask.dtype == torch.float32 else attn_mask.float()
    return AggregateFunction.apply(x, attn, pad_mask, mode)


def temporal_aggregate_forward(x: torch.Tensor, pad_mask: Optional[torch.Tensor] = None,
                               attn_mask: Optional[torch.Tensor] = None, mode: str = "mean") -> torch.Tensor:
    """``TemporalAggregator(mode).forward(x, pad_mask, attn_mask)`` (temporal_aggregator.py:14-77).

    x[B,T,C,H,W] float32/bfloat16, attn_mask[h,B,T,ha,wa] (float32 used as is), pad_mask[B,T] bool.
    Returns out[B,C,H,W] in x's dtype.  One kernel launch; no host synchronisation.
    """
    if mode not in _AGG_MODES:
        raise ValueError(f"unknown aggregation mode {mode!r}")
    if x.dim() != 5:
        raise RuntimeError(f"crop2seg_b200: x must be [B,T,C,H,W], got {tuple(x.shape)}")
    _require_cuda(x, "x")
    x = x.contiguous()
    b, t, c, h, w = x.shape
    desc = _lib.AggDesc(B=b, T=t, C=c, H=h, W=w, n_heads=1, ha=1, wa=1, mode=_AGG_MODES[mode],
                        dtype=_dtype_code(x, "x"))
    attn = None
    if mode != "mean":
        if attn_mask is None:
            raise RuntimeError(f"crop2seg_b200: mode {mode!r} needs attn_mask")
        if attn_mask.dim() != 5 or attn_mask.shape[1] != b or attn_mask.shape[2] != t:
            raise RuntimeError(
                f"crop2seg_b200: attn_mask must be [h,{b},{t},ha,wa], got {tuple(attn_mask.shape)}")
        attn = attn_mask.to(device=x.device, dtype=torch.float32).contiguous()
        desc.n_heads, desc.ha, desc.wa = attn.shape[0], attn.shape[3], attn.shape[4]
    pad = _mask_u8(pad_mask, b, t, x.device)
    out = torch.empty((b, c, h, w), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws_bytes = lib.c2s_agg_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        status = lib.c2s_agg_forward(ctypes.byref(desc), x.data_ptr(), _ptr(attn), _ptr(pad), out.data_ptr(),
                                     _ptr(ws), ws_bytes, _stream(x.device))
    _lib.check(status, "c2s_agg_forward")
    return out


def temporal_aggregate_skip_conv(x: torch.Tensor, pad_mask: Optional[torch.Tensor], attn_mask: torch.Tensor,
                                 conv_weight: torch.Tensor, conv_bias: Optional[torch.Tensor],
                                 bn_weight: Optional[torch.Tensor], bn_bias: Optional[torch.Tensor],
                                 bn_running_mean: torch.Tensor, bn_running_var: torch.Tensor, bn_eps: float = 1e-5
                                 ) -> torch.Tensor:
    """``relu(BatchNorm2d_eval(Conv2d_1x1(TemporalAggregator('att_group')(x, pad_mask, attn_mask))))`` in one kernel:
    the aggregation of utae.py:225-227 followed by ``UpConvBlock.skip_conv`` (conv.py:378-382, 408), the skip map never
    leaving the chip (``c2s_agg_skipconv_forward``).  Inference only; bf16 x[B,T,64,H,W], 16 heads, x2/x4/x8
    up-sampling, H*W % 128 == 0 -- other shapes raise (use the two modules separately)."""
    if x.dim() != 5:
        raise RuntimeError(f"crop2seg_b200: x must be [B,T,C,H,W], got {tuple(x.shape)}")
    _require_cuda(x, "x")
    x = x.contiguous()
    b, t, c, h, w = x.shape
    if attn_mask is None or attn_mask.dim() != 5 or attn_mask.shape[1] != b or attn_mask.shape[2] != t:
        raise RuntimeError(f"crop2seg_b200: attn_mask must be [h,{b},{t},ha,wa]")
    if tuple(conv_weight.shape[:2]) != (c, c) or conv_weight.numel() != c * c:
        raise RuntimeError(f"crop2seg_b200: skip convolution must be 1x1 {c}->{c}, got {tuple(conv_weight.shape)}")
    attn = attn_mask.to(device=x.device, dtype=torch.float32).contiguous()
    desc = _lib.AggDesc(B=b, T=t, C=c, H=h, W=w, n_heads=attn.shape[0], ha=attn.shape[3], wa=attn.shape[4],
                        mode=_AGG_MODES["att_group"], dtype=_dtype_code(x, "x"))
    pad = _mask_u8(pad_mask, b, t, x.device)
    keep = []
    params = _lib.SkipConvParams(
        conv_weight=_f32(conv_weight, x.device, keep), conv_bias=_f32(conv_bias, x.device, keep),
        bn_weight=_f32(bn_weight, x.device, keep), bn_bias=_f32(bn_bias, x.device, keep),
        bn_running_mean=_f32(bn_running_mean, x.device, keep), bn_running_var=_f32(bn_running_var, x.device, keep),
        bn_eps=float(bn_eps))
    out = torch.empty((b, c, h, w), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws_bytes = lib.c2s_agg_skipconv_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        status = lib.c2s_agg_skipconv_forward(ctypes.byref(desc), x.data_ptr(), attn.data_ptr(), _ptr(pad),
                                              ctypes.byref(params), out.data_ptr(), ws.data_ptr(), ws_bytes,
                                              _stream(x.device))
    _lib.check(status, "c2s_agg_skipconv_forward")
    return out


def temporal_aggregate_backward(x: torch.Tensor, pad_mask: Optional[torch.Tensor], attn_mask: Optional[torch.Tensor],
                                grad_out: torch.Tensor, mode: str, need_x: bool = True, need_attn: bool = True
                                ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Backward of :func:`temporal_aggregate`: ``(grad_x | None, grad_attn | None)`` (``c2s_agg_backward``).

    grad_attn is accumulated with float atomics (low bits depend on the execution order, like the reference's own
    backward, train.py:623-626)."""
    _require_cuda(x, "x")
    x = x.contiguous()
    b, t, c, h, w = x.shape
    desc = _lib.AggDesc(B=b, T=t, C=c, H=h, W=w, n_heads=1, ha=1, wa=1, mode=_AGG_MODES[mode],
                        dtype=_dtype_code(x, "x"))
    attn = None
    if mode != "mean":
        attn = attn_mask.to(device=x.device, dtype=torch.float32).contiguous()
        desc.n_heads, desc.ha, desc.wa = attn.shape[0], attn.shape[3], attn.shape[4]
    else:
        need_attn = False
    gout = grad_out.to(dtype=x.dtype).contiguous()
    pad = _mask_u8(pad_mask, b, t, x.device)
    gx = torch.empty_like(x) if need_x else None
    gattn = torch.zeros_like(attn) if need_attn else None
    lib = _lib.load()
    with torch.cuda.device(x.device):
        ws_bytes = lib.c2s_agg_backward_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
        status = lib.c2s_agg_backward(ctypes.byref(desc), x.data_ptr(), _ptr(attn), _ptr(pad), gout.data_ptr(),
                                      _ptr(gx), _ptr(gattn), _ptr(ws), ws_bytes, _stream(x.device))
    _lib.check(status, "c2s_agg_backward")
    return gx, gattn


def _f32(t: Optional[torch.Tensor], device, keep) -> Optional[int]:
    """Device pointer of a float32, contiguous view of a parameter (kept alive in ``keep``)."""
    if t is None:
        return None
    v = t.detach()
    if v.device != device or v.dtype != torch.float32 or not v.is_contiguous():
        v = v.to(device=device, dtype=torch.float32).contiguous()
    keep.append(v)
    return v.data_ptr()


def ltae_forward(x: torch.Tensor, positions: Optional[torch.Tensor], pad_mask: Optional[torch.Tensor],
                 params: Dict[str, Optional[torch.Tensor]], *, n_head: int, d_k: int, d_model: int,
                 has_inconv: bool, c_out: int, pe_mode: int, pe_abs: bool = False, attn_only: bool = False,
                 need_attn: bool = True, zero_padded: bool = False, bn_batch_stats: bool = False,
                 gn_eps: float = 1e-5, bn_eps: float = 1e-5, attn_keep: Optional[torch.Tensor] = None,
                 attn_drop_p: float = 0.0, mlp_keep: Optional[torch.Tensor] = None, mlp_drop_p: float = 0.0,
                 folded_cache: Optional[dict] = None, folded_key=None, save_o: bool = False, save_y: bool = False
                 ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor], Optional[Tuple[torch.Tensor, torch.Tensor]]]:
    """Fused ``LTAE.forward`` / ``LTAE4WTAE.forward`` (tae.py:451-504, 589-635).

    ``params`` maps the field names of ``c2s_ltae_params`` to tensors (or None).
    ``folded_cache`` (a dict owned by the caller) with ``folded_key`` (anything that changes whenever a parameter
    does) keeps the workspace between calls: while key, shapes and flags repeat, the weight-only preparation kernels
    are skipped (``C2S_LTAE_REUSE_FOLDED``).  Workspaces above 64 MiB are never kept.
    Returns ``(out[B,c_out,H,W] | None, attn[h,B,T,H,W] | None, (batch_mean, batch_var) | None)``; with ``save_o`` a
    fourth item, the float32 rows ``o[B*H*W, d_model]`` that enter the MLP (needed by :func:`ltae_backward`'s caller).
    """
    if x.dim() != 5:
        raise RuntimeError(f"crop2seg_b200: x must be [B,T,C,H,W], got {tuple(x.shape)}")
    _require_cuda(x, "x")
    x = x.contiguous()
    dev = x.device
    b, t, c, h, w = x.shape
    flags = 0
    if attn_only:
        flags |= _lib.LTAE_ATTN_ONLY
    if not need_attn:
        flags |= _lib.LTAE_SKIP_ATTN_STORE
    if zero_padded:
        flags |= _lib.LTAE_ZERO_PADDED
    if bn_batch_stats and not attn_only:
        flags |= _lib.LTAE_BN_BATCH_STATS
    pos = None
    pos_dtype = 0
    if pe_mode != _lib.PE_NONE:
        if positions is None:
            raise RuntimeError("crop2seg_b200: batch_positions is required when positional encoding is enabled")
        want = (b, t, 2) if pe_abs else (b, t)
        if tuple(positions.shape) != want:
            raise RuntimeError(f"crop2seg_b200: batch_positions has shape {tuple(positions.shape)}, expected {want}")
        pos = positions.to(device=dev)
        if pos.dtype.is_floating_point:
            pos = pos.to(torch.float32)
            pos_dtype = 1
        else:
            pos = pos.to(torch.int64)
        pos = pos.contiguous()
    desc = _lib.LtaeDesc(B=b, T=t, C=c, H=h, W=w, n_head=n_head, d_k=d_k, d_model=d_model,
                         c_out=0 if attn_only else c_out, has_inconv=int(has_inconv), pe_mode=pe_mode,
                         pe_abs=int(pe_abs), pos_dtype=pos_dtype, dtype=_dtype_code(x, "x"), flags=flags,
                         gn_eps=gn_eps, bn_eps=bn_eps, attn_keep_scale=1.0 / (1.0 - attn_drop_p),
                         mlp_keep_scale=1.0 / (1.0 - mlp_drop_p))
    keep = []
    cparams = _lib.LtaeParams(**{k: _f32(params.get(k), dev, keep) for k in _lib.LTAE_PARAM_FIELDS})
    if attn_keep is not None:  # uint8 [h,B,T,H,W]: dropout mask of the attention (tae.py:837)
        if tuple(attn_keep.shape) != (n_head, b, t, h, w):
            raise RuntimeError(f"crop2seg_b200: attn_keep has shape {tuple(attn_keep.shape)}")
        m = attn_keep.to(device=dev, dtype=torch.uint8).contiguous()
        keep.append(m)
        cparams.attn_keep = m.data_ptr()
    if mlp_keep is not None and not attn_only:  # uint8 [B,c_out,H,W]: dropout mask after the ReLU (tae.py:448)
        if tuple(mlp_keep.shape) != (b, c_out, h, w):
            raise RuntimeError(f"crop2seg_b200: mlp_keep has shape {tuple(mlp_keep.shape)}")
        m = mlp_keep.to(device=dev, dtype=torch.uint8).contiguous()
        keep.append(m)
        cparams.mlp_keep = m.data_ptr()
    o_rows = y_rows = None
    if save_o and not attn_only:
        o_rows = torch.empty((b * h * w, d_model), dtype=torch.float32, device=dev)
        cparams.save_o = o_rows.data_ptr()
    if save_y and not attn_only and bn_batch_stats:  # the pre-BatchNorm rows of the training-mode forward, for its backward
        y_rows = torch.empty((b * h * w, c_out), dtype=torch.float32, device=dev)
        cparams.save_y = y_rows.data_ptr()
    pad = _mask_u8(pad_mask, b, t, dev)
    out = None if attn_only else torch.empty((b, c_out, h, w), dtype=x.dtype, device=dev)
    attn = torch.empty((n_head, b, t, h, w), dtype=torch.float32, device=dev) if need_attn else None
    stats = None
    if flags & _lib.LTAE_BN_BATCH_STATS:
        stats = (torch.empty(c_out, dtype=torch.float32, device=dev),
                 torch.empty(c_out, dtype=torch.float32, device=dev))
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws_bytes = lib.c2s_ltae_workspace_bytes(ctypes.byref(desc))
        ws = None
        # Under CUDA-graph capture the cache is bypassed: a captured graph bakes the workspace pointer, and a cached
        # workspace can be evicted (another shape) or go stale (weights updated) while the graph lives on.  A captured
        # call allocates its workspace inside the capture (the graph's private pool keeps it alive) and folds the
        # weights on every replay.
        capturing = torch.cuda.is_current_stream_capturing()
        if folded_cache is not None and ws_bytes <= _FOLDED_CACHE_MAX_BYTES and not capturing:
            key = (folded_key, dev, b, t, c, h, w, n_head, d_k, d_model, desc.c_out, int(has_inconv), pe_mode,
                   int(pe_abs), desc.dtype, flags, float(gn_eps), float(bn_eps))
            if folded_cache.get("key") == key and folded_cache.get("ws") is not None:
                ws = folded_cache["ws"]
                desc.flags = flags | _lib.LTAE_REUSE_FOLDED
            else:
                ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
                folded_cache["key"], folded_cache["ws"] = key, ws
        if ws is None:
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        status = lib.c2s_ltae_forward(ctypes.byref(desc), ctypes.byref(cparams), x.data_ptr(), _ptr(pos), _ptr(pad),
                                      _ptr(out), _ptr(attn), _ptr(stats[0]) if stats else None,
                                      _ptr(stats[1]) if stats else None, ws.data_ptr(), ws_bytes, _stream(dev))
    _lib.check(status, "c2s_ltae_forward")
    if save_y:
        return out, attn, stats, o_rows, y_rows
    if save_o:
        return out, attn, stats, o_rows
    return out, attn, stats


def ltae_rows_forward(o_rows: torch.Tensor, params: Dict[str, Optional[torch.Tensor]], *, batch: int, height: int,
                      width: int, n_head: int, d_model: int, c_out: int, dtype: torch.dtype, gn_eps: float = 1e-5,
                      bn_eps: float = 1e-5) -> torch.Tensor:
    """``c2s_ltae_rows_forward``: the rows behind the attention of an encoder with n > 1 learned queries, eval mode
    (tae.py:486-499): ``o_rows`` [n, B*H*W, d_model] (float32, the ``save_o`` rows of one ``ltae_forward`` per query)
    -> ``out`` [B, n, c_out, H, W]; the output GroupNorm runs over the channels of a group and the n queries."""
    _require_cuda(o_rows, "o_rows")
    dev = o_rows.device
    n_q = o_rows.shape[0]
    if o_rows.dim() != 3 or tuple(o_rows.shape[1:]) != (batch * height * width, d_model):
        raise RuntimeError(f"crop2seg_b200: o_rows has shape {tuple(o_rows.shape)}")
    o = o_rows.to(torch.float32).contiguous()
    out = torch.empty((batch, n_q, c_out, height, width), dtype=dtype, device=dev)
    desc = _lib.LtaeDesc(B=batch, T=1, C=n_head, H=height, W=width, n_head=n_head, d_k=1, d_model=d_model, c_out=c_out,
                         has_inconv=1, pe_mode=_lib.PE_NONE, pe_abs=0, pos_dtype=0, dtype=_dtype_code(out, "out"), flags=0,
                         gn_eps=gn_eps, bn_eps=bn_eps, attn_keep_scale=1.0, mlp_keep_scale=1.0)
    keep = []
    cparams = _lib.LtaeParams(**{k: _f32(params.get(k), dev, keep) for k in _lib.LTAE_PARAM_FIELDS})
    with torch.cuda.device(dev):
        status = _lib.load().c2s_ltae_rows_forward(ctypes.byref(desc), ctypes.byref(cparams), o.data_ptr(), n_q,
                                                   out.data_ptr(), _stream(dev))
    _lib.check(status, "c2s_ltae_rows_forward")
    return out


def ltae_backward(x: torch.Tensor, positions: Optional[torch.Tensor], pad_mask: Optional[torch.Tensor],
                  params: Dict[str, Optional[torch.Tensor]], grad_o: Optional[torch.Tensor],
                  grad_attn: Optional[torch.Tensor], *, n_head: int, d_k: int, d_model: int, has_inconv: bool,
                  c_out: int, pe_mode: int, pe_abs: bool = False, attn_only: bool = False, zero_padded: bool = False,
                  gn_eps: float = 1e-5, attn_keep: Optional[torch.Tensor] = None, attn_drop_p: float = 0.0,
                  need_grad_pe: bool = False) -> Dict[str, Optional[torch.Tensor]]:
    """``c2s_ltae_backward``: everything of the L-TAE backward that touches the [B*H*W, T, C] features.

    ``grad_o`` [B*H*W, d_model] is the gradient w.r.t. the rows entering the MLP, ``grad_attn`` [h,B,T,H,W] the one
    w.r.t. the returned attention.  Returns ``grad_x`` and the gradients of the folded quantities (``grad_u`` [C,16],
    ``grad_cpos`` [B,T,16], ``grad_gamma`` / ``grad_beta`` [C] direct terms, ``zn_rows`` [B*H*W,h,C], ``sa_rows``
    [B*H*W,16], ``grad_pe`` [B,T,d_model] | None); see ``include/crop2seg_b200.h``."""
    _require_cuda(x, "x")
    x = x.contiguous()
    dev = x.device
    b, t, c, h, w = x.shape
    n = b * h * w
    flags = (_lib.LTAE_ATTN_ONLY if attn_only else 0) | (_lib.LTAE_ZERO_PADDED if zero_padded else 0)
    pos, pos_dtype = None, 0
    if pe_mode != _lib.PE_NONE:
        pos = positions.to(device=dev)
        if pos.dtype.is_floating_point:
            pos, pos_dtype = pos.to(torch.float32), 1
        else:
            pos = pos.to(torch.int64)
        pos = pos.contiguous()
    desc = _lib.LtaeDesc(B=b, T=t, C=c, H=h, W=w, n_head=n_head, d_k=d_k, d_model=d_model,
                         c_out=0 if attn_only else c_out, has_inconv=int(has_inconv), pe_mode=pe_mode,
                         pe_abs=int(pe_abs), pos_dtype=pos_dtype, dtype=_dtype_code(x, "x"), flags=flags,
                         gn_eps=gn_eps, bn_eps=1e-5, attn_keep_scale=1.0 / (1.0 - attn_drop_p), mlp_keep_scale=1.0)
    keep = []
    cparams = _lib.LtaeParams(**{k: _f32(params.get(k), dev, keep) for k in _lib.LTAE_PARAM_FIELDS})
    if attn_keep is not None:
        m = attn_keep.to(device=dev, dtype=torch.uint8).contiguous()
        keep.append(m)
        cparams.attn_keep = m.data_ptr()
    f32 = dict(dtype=torch.float32, device=dev)
    # the accumulated (atomic) outputs are slices of ONE zero-filled buffer: one memset instead of one per tensor
    want_pe = (not attn_only) and need_grad_pe and pe_mode != _lib.PE_NONE
    sizes = [c * 16, b * t * 16] + ([c, c] if not attn_only else []) + ([b * t * d_model] if want_pe else [])
    offs = [0]
    for sz in sizes:
        offs.append(offs[-1] + (sz + 3) // 4 * 4)  # 16-byte aligned slices
    zeros = torch.zeros(offs[-1], dtype=torch.float32, device=dev)
    res = {
        "grad_x": torch.empty_like(x), "grad_u": zeros[offs[0]:offs[0] + c * 16].view(c, 16),
        "grad_cpos": zeros[offs[1]:offs[1] + b * t * 16].view(b, t, 16),
        "grad_gamma": None, "grad_beta": None, "zn_rows": None, "sa_rows": None, "grad_pe": None,
    }
    io = _lib.LtaeBwdIo()
    if not attn_only:
        if grad_o is None or tuple(grad_o.shape) != (n, d_model):
            raise RuntimeError(f"crop2seg_b200: grad_o must be [{n},{d_model}]")
        go = grad_o.to(**f32).contiguous()
        keep.append(go)
        io.grad_o = go.data_ptr()
        res["grad_gamma"], res["grad_beta"] = zeros[offs[2]:offs[2] + c], zeros[offs[3]:offs[3] + c]
        res["zn_rows"], res["sa_rows"] = torch.empty((n, n_head, c), **f32), torch.empty((n, 16), **f32)
        if want_pe:
            res["grad_pe"] = zeros[offs[4]:offs[4] + b * t * d_model].view(b, t, d_model)
    if grad_attn is not None:
        ga = grad_attn.to(**f32).contiguous()
        if tuple(ga.shape) != (n_head, b, t, h, w):
            raise RuntimeError(f"crop2seg_b200: grad_attn has shape {tuple(ga.shape)}")
        keep.append(ga)
        io.grad_attn = ga.data_ptr()
    for k in ("grad_x", "grad_u", "grad_cpos", "grad_gamma", "grad_beta", "zn_rows", "sa_rows", "grad_pe"):
        if res[k] is not None:
            setattr(io, k, res[k].data_ptr())
    pad = _mask_u8(pad_mask, b, t, dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws_bytes = lib.c2s_ltae_backward_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        status = lib.c2s_ltae_backward(ctypes.byref(desc), ctypes.byref(cparams), x.data_ptr(), _ptr(pos), _ptr(pad),
                                       ctypes.byref(io), ws.data_ptr(), ws_bytes, _stream(dev))
    _lib.check(status, "c2s_ltae_backward")
    res["_fold"] = (desc, cparams, keep, ws, ws_bytes)  # c2s_ltae_fold_backward reads the folded tensors of this workspace
    return res


def ltae_fold_backward(res: Dict[str, torch.Tensor], want: Dict[str, bool], shapes: Dict[str, tuple],
                       grad_pe_through_scores: bool = False) -> Dict[str, torch.Tensor]:
    """``c2s_ltae_fold_backward`` on the result of :func:`ltae_backward`: gradients of ``in_norm_weight/bias``,
    ``inconv_weight/bias`` (folded part), ``query``, ``key_weight``, ``key_bias`` -- the ones ``want`` names --
    as float32 tensors of ``shapes``.  ``grad_pe_through_scores``: also add grad_cpos . qk to ``res['grad_pe']``."""
    desc, cparams, keep, ws, ws_bytes = res["_fold"]
    dev = res["grad_u"].device
    io = _lib.LtaeFoldBwdIo(grad_u=res["grad_u"].data_ptr(), grad_cpos=res["grad_cpos"].data_ptr(),
                            grad_gamma_direct=_ptr(res.get("grad_gamma")), grad_beta_direct=_ptr(res.get("grad_beta")))
    out = {}
    for k in ("in_norm_weight", "in_norm_bias", "inconv_weight", "inconv_bias", "query", "key_weight", "key_bias"):
        if want.get(k):
            out[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
            setattr(io, "grad_" + k, out[k].data_ptr())
    if grad_pe_through_scores and res.get("grad_pe") is not None:
        io.grad_pe = res["grad_pe"].data_ptr()
    with torch.cuda.device(dev):
        status = _lib.load().c2s_ltae_fold_backward(ctypes.byref(desc), ctypes.byref(cparams), ctypes.byref(io), ws.data_ptr(),
                                                    ws_bytes, _stream(dev))
    _lib.check(status, "c2s_ltae_fold_backward")
    return out


def ltae_mlp_backward(o_rows: torch.Tensor, grad_out: torch.Tensor, params: Dict[str, Optional[torch.Tensor]],
                      bn_mean: torch.Tensor, bn_var: torch.Tensor, *, n_head: int, d_model: int, c_out: int,
                      bn_batch_stats: bool, gn_eps: float = 1e-5, bn_eps: float = 1e-5,
                      mlp_keep: Optional[torch.Tensor] = None, mlp_drop_p: float = 0.0,
                      y_rows: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """``c2s_ltae_mlp_backward``: backward of mlp.0 / mlp.2 / ReLU / dropout / out_norm on the pixel rows.

    ``o_rows`` [B*H*W, d_model] are the rows saved by the forward, ``grad_out`` [B, c_out, H, W] the incoming gradient,
    ``bn_mean`` / ``bn_var`` the statistics the forward normalised with, ``y_rows`` [B*H*W, c_out] the pre-BatchNorm rows
    the training-mode forward saved (``save_y``; recomputed from ``o_rows`` when None).  Returns ``grad_o`` and the
    parameter gradients keyed like ``c2s_ltae_params``."""
    _require_cuda(grad_out, "grad_out")
    dev = grad_out.device
    b, co, h, w = grad_out.shape
    n = b * h * w
    if co != c_out or tuple(o_rows.shape) != (n, d_model):
        raise RuntimeError(f"crop2seg_b200: grad_out {tuple(grad_out.shape)} / o_rows {tuple(o_rows.shape)} do not match")
    g = grad_out.contiguous()
    flags = _lib.LTAE_BN_BATCH_STATS if bn_batch_stats else 0
    desc = _lib.LtaeDesc(B=b, T=1, C=n_head, H=h, W=w, n_head=n_head, d_k=1, d_model=d_model, c_out=c_out, has_inconv=1,
                         pe_mode=_lib.PE_NONE, pe_abs=0, pos_dtype=0, dtype=_dtype_code(g, "grad_out"), flags=flags,
                         gn_eps=gn_eps, bn_eps=bn_eps, attn_keep_scale=1.0, mlp_keep_scale=1.0 / (1.0 - mlp_drop_p))
    keep = []
    cparams = _lib.LtaeParams(**{k: _f32(params.get(k), dev, keep) for k in _lib.LTAE_PARAM_FIELDS})
    if mlp_keep is not None:
        m = mlp_keep.to(device=dev, dtype=torch.uint8).contiguous()
        keep.append(m)
        cparams.mlp_keep = m.data_ptr()
    f32 = dict(dtype=torch.float32, device=dev)
    cpad = (c_out + 3) // 4 * 4
    zeros = torch.zeros(c_out * d_model + 5 * cpad, dtype=torch.float32, device=dev)  # one memset for the six accumulators
    res = {"grad_o": torch.empty((n, d_model), **f32), "mlp_weight": zeros[:c_out * d_model].view(c_out, d_model)}
    for i, k in enumerate(("mlp_bias", "bn_weight", "bn_bias", "out_norm_weight", "out_norm_bias")):
        res[k] = zeros[c_out * d_model + i * cpad:c_out * d_model + i * cpad + c_out]
    o = o_rows.to(**f32).contiguous()
    mean, var = bn_mean.to(**f32).contiguous(), bn_var.to(**f32).contiguous()
    yr = None
    if y_rows is not None:
        if tuple(y_rows.shape) != (n, c_out):
            raise RuntimeError(f"crop2seg_b200: y_rows has shape {tuple(y_rows.shape)}, expected {(n, c_out)}")
        yr = y_rows.to(**f32).contiguous()
    io = _lib.LtaeMlpBwdIo(o_rows=o.data_ptr(), y_rows=_ptr(yr), grad_out=g.data_ptr(), bn_mean=mean.data_ptr(), bn_var=var.data_ptr(),
                           grad_o=res["grad_o"].data_ptr(), grad_mlp_weight=res["mlp_weight"].data_ptr(),
                           grad_mlp_bias=res["mlp_bias"].data_ptr(), grad_bn_weight=res["bn_weight"].data_ptr(),
                           grad_bn_bias=res["bn_bias"].data_ptr(), grad_out_norm_weight=res["out_norm_weight"].data_ptr(),
                           grad_out_norm_bias=res["out_norm_bias"].data_ptr())
    lib = _lib.load()
    with torch.cuda.device(dev):
        ws_bytes = lib.c2s_ltae_mlp_backward_workspace_bytes(ctypes.byref(desc))
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        status = lib.c2s_ltae_mlp_backward(ctypes.byref(desc), ctypes.byref(cparams), ctypes.byref(io), ws.data_ptr(),
                                           ws_bytes, _stream(dev))
    _lib.check(status, "c2s_ltae_mlp_backward")
    return res


def ltae_inconv_grad(grad_o: torch.Tensor, zn_rows: torch.Tensor, sa_rows: Optional[torch.Tensor],
                     grad_weight: torch.Tensor, grad_bias: Optional[torch.Tensor], n_head: int) -> None:
    """``c2s_ltae_inconv_grad``: add the direct term of the in-projection gradient to ``grad_weight`` [d_model, C]
    (and ``grad_bias`` [d_model]) from ``grad_o`` [N, d_model], ``zn_rows`` [N, n_head, C], ``sa_rows`` [N, 16]."""
    dev = grad_o.device
    n, d_model = grad_o.shape
    c = zn_rows.shape[-1]
    for name, t in (("grad_o", grad_o), ("zn_rows", zn_rows), ("grad_weight", grad_weight)):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
            raise RuntimeError(f"crop2seg_b200: {name} must be a contiguous float32 tensor on {dev}")
    if grad_bias is not None and (sa_rows is None or not sa_rows.is_contiguous() or not grad_bias.is_contiguous()):
        raise RuntimeError("crop2seg_b200: the bias gradient needs contiguous sa_rows / grad_bias")
    lib = _lib.load()
    with torch.cuda.device(dev):
        status = lib.c2s_ltae_inconv_grad(grad_o.data_ptr(), zn_rows.data_ptr(), _ptr(sa_rows), grad_weight.data_ptr(),
                                          _ptr(grad_bias), n, n_head, d_model, c, _stream(dev))
    _lib.check(status, "c2s_ltae_inconv_grad")
