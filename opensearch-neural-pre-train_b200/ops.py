"""Torch-tensor wrappers and autograd Functions over the C-ABI (``include/splade_b200.h``).

PyTorch is plumbing here: it allocates every buffer (inputs, outputs, workspaces), supplies the current CUDA
stream and records the autograd graph.  All arithmetic happens in ``libsplade_b200.so``.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, load_library


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"splade_b200: `{name}` is on {t.device}; this path runs only on an sm_100 CUDA device "
            f"(no CPU fallback)."
        )


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _ws(nbytes: int, device) -> Tuple[Optional[torch.Tensor], int]:
    if nbytes <= 0:
        return None, 0
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return buf, buf.data_ptr()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------
# raw ops
# ------------------------------------------------------------------------------------------------
def head_forward(hidden: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, mask: torch.Tensor,
                 want_token_weights: bool = True, pack_mode: Optional[int] = None,
                 plan_out: Optional[torch.Tensor] = None):
    """hidden [B,S,H] bf16, weight [V,H] bf16, bias [V] fp32, mask [B,S] int32 ->
    (rep [B,V] fp32, argmax [B,V] int32, token_weights [B,S] fp32 | None).

    ``pack_mode``: None = the library default (env SPLADE_B200_PACK, "auto": a device-side plan picks the padded
    or the packed variable-length tiling without a host sync); -1 auto, 0 padded, 1 packed (``splade_head_fwd_mode``).
    ``plan_out``: optional int32[4] CUDA tensor receiving {tiling used, real tokens, packed tiles, last-tile N}."""
    lib = load_library()
    for n, t in (("hidden", hidden), ("weight", weight), ("bias", bias), ("mask", mask)):
        _require_cuda(t, n)
    assert hidden.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
    assert bias.dtype == torch.float32 and mask.dtype == torch.int32
    assert hidden.is_contiguous() and weight.is_contiguous() and bias.is_contiguous() and mask.is_contiguous()
    B, S, H = hidden.shape
    V = weight.shape[0]
    assert weight.shape[1] == H and bias.shape[0] == V and tuple(mask.shape) == (B, S)
    dev = hidden.device
    with torch.cuda.device(dev):
        rep = torch.empty((B, V), dtype=torch.float32, device=dev)
        argmax = torch.empty((B, V), dtype=torch.int32, device=dev)
        tokw = torch.empty((B, S), dtype=torch.float32, device=dev) if want_token_weights else None
        op = _lib.OP_HEAD_FWD_TOKW if want_token_weights else _lib.OP_HEAD_FWD
        nbytes = lib.splade_workspace_bytes(op, B, S, H, V, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        if pack_mode is None and plan_out is None:
            check(lib.splade_head_fwd(hidden.data_ptr(), weight.data_ptr(), bias.data_ptr(), mask.data_ptr(),
                                      B, S, H, V, rep.data_ptr(), argmax.data_ptr(), _ptr(tokw), ws_ptr, nbytes,
                                      _stream(hidden)), "splade_head_fwd")
        else:
            if plan_out is not None:
                assert plan_out.is_cuda and plan_out.dtype == torch.int32 and plan_out.numel() >= 4
            check(lib.splade_head_fwd_mode(hidden.data_ptr(), weight.data_ptr(), bias.data_ptr(), mask.data_ptr(),
                                           B, S, H, V, rep.data_ptr(), argmax.data_ptr(), _ptr(tokw), ws_ptr,
                                           nbytes, -1 if pack_mode is None else int(pack_mode), _ptr(plan_out),
                                           _stream(hidden)), "splade_head_fwd_mode")
    return rep, argmax, tokw


class HostEncodeSlot:
    """Device staging buffers for one in-flight host-buffer encode (``splade_head_fwd_host``)."""

    def __init__(self, B: int, S: int, H: int, V: int, device):
        lib = load_library()
        self.shape = (B, S, H, V)
        self.hidden = torch.empty((B, S, H), dtype=torch.bfloat16, device=device)
        self.mask = torch.empty((B, S), dtype=torch.int32, device=device)
        self.rep = torch.empty((B, V), dtype=torch.float32, device=device)
        self.argmax = torch.empty((B, V), dtype=torch.int32, device=device)
        self.ws_bytes = lib.splade_workspace_bytes(_lib.OP_HEAD_FWD, B, S, H, V, 0)
        self.ws = torch.empty(max(self.ws_bytes, 1), dtype=torch.uint8, device=device)


def head_forward_host(hidden_host: torch.Tensor, mask_host: torch.Tensor, weight: torch.Tensor,
                      bias: torch.Tensor, slot: HostEncodeSlot, rep_host: torch.Tensor,
                      argmax_host: Optional[torch.Tensor] = None, stream: Optional[torch.cuda.Stream] = None):
    """Host buffers in, host buffers out (pinned CPU tensors): H2D -> fused head -> D2H enqueued on `stream`
    through ONE C-ABI call; returns immediately (synchronise the stream before reading ``rep_host``)."""
    lib = load_library()
    assert not hidden_host.is_cuda and hidden_host.dtype == torch.bfloat16 and hidden_host.is_contiguous()
    assert not mask_host.is_cuda and mask_host.dtype == torch.int32 and not rep_host.is_cuda
    _require_cuda(weight, "weight")
    B, S, H, V = slot.shape
    assert tuple(hidden_host.shape) == (B, S, H) and tuple(rep_host.shape) == (B, V)
    st = stream if stream is not None else torch.cuda.current_stream(weight.device)
    with torch.cuda.device(weight.device):
        check(lib.splade_head_fwd_host(hidden_host.data_ptr(), mask_host.data_ptr(), weight.data_ptr(),
                                       bias.data_ptr(), B, S, H, V, slot.hidden.data_ptr(), slot.mask.data_ptr(),
                                       slot.rep.data_ptr(), slot.argmax.data_ptr(), rep_host.data_ptr(),
                                       _ptr(argmax_host), slot.ws.data_ptr(), slot.ws_bytes, st.cuda_stream),
              "splade_head_fwd_host")


def head_backward(g_rep: torch.Tensor, rep: torch.Tensor, argmax: torch.Tensor, hidden: torch.Tensor,
                  weight: torch.Tensor):
    """-> (d_hidden [B,S,H] fp32, dW [V,H] fp32, dbias [V] fp32)"""
    lib = load_library()
    B, S, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    g_rep = _f32c(g_rep)
    with torch.cuda.device(dev):
        d_hidden = torch.empty((B, S, H), dtype=torch.float32, device=dev)
        dW = torch.empty((V, H), dtype=torch.float32, device=dev)
        dbias = torch.empty((V,), dtype=torch.float32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_HEAD_BWD, B, S, H, V, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_head_bwd(g_rep.data_ptr(), rep.data_ptr(), argmax.data_ptr(), hidden.data_ptr(),
                                  weight.data_ptr(), B, S, H, V, d_hidden.data_ptr(), dW.data_ptr(),
                                  dbias.data_ptr(), ws_ptr, nbytes, _stream(hidden)), "splade_head_bwd")
    return d_hidden, dW, dbias


def _segment_struct(g_rep, rep, argmax, hidden, d_hidden, active=None):
    B, S, _ = hidden.shape
    return _lib.HeadBwdSegment(g_rep=g_rep.data_ptr(), rep=rep.data_ptr(), argmax=argmax.data_ptr(),
                               hidden=hidden.data_ptr(), d_hidden=_ptr(d_hidden), B=B, S=S, active=_ptr(active))


def rep_bitmap(rep: torch.Tensor) -> torch.Tensor:
    """rep [B,V] fp32 -> int32 [B, ceil(V/32)], bit j of word w = (rep[b, 32w+j] > 0) (``splade_rep_bitmap``).
    The backward kernels are driven by it; computed right after the forward, while rep is L2-hot."""
    lib = load_library()
    _require_cuda(rep, "rep")
    assert rep.dtype == torch.float32 and rep.is_contiguous()
    B, V = rep.shape
    with torch.cuda.device(rep.device):
        bitmap = torch.empty((B, (V + 31) // 32), dtype=torch.int32, device=rep.device)
        check(lib.splade_rep_bitmap(rep.data_ptr(), B, V, bitmap.data_ptr(), _stream(rep)), "splade_rep_bitmap")
    return bitmap


def head_backward_dhidden(g_rep: torch.Tensor, rep: torch.Tensor, argmax: torch.Tensor, hidden: torch.Tensor,
                          weight: torch.Tensor, want_d_hidden: bool = True, active: Optional[torch.Tensor] = None):
    """Phase 1 of the multi-segment backward for ONE head call (``splade_head_bwd_dhidden``):
    -> (d_hidden [B,S,H] fp32 | None, activity bitmap int32 [B, ceil(V/32)] for ``head_backward_dw``).
    ``active``: the bitmap ``rep_bitmap(rep)`` made after the forward; computed here when missing."""
    lib = load_library()
    B, S, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    assert g_rep.dtype == torch.float32 and g_rep.is_contiguous()
    with torch.cuda.device(dev):
        d_hidden = torch.empty((B, S, H), dtype=torch.float32, device=dev) if want_d_hidden else None
        bitmap = active if active is not None else torch.empty((B, (V + 31) // 32), dtype=torch.int32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_HEAD_BWD_DHIDDEN, B, S, H, V, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        seg = _segment_struct(g_rep, rep, argmax, hidden, d_hidden, active)
        check(lib.splade_head_bwd_dhidden(ctypes.byref(seg), weight.data_ptr(), H, V, bitmap.data_ptr(), ws_ptr,
                                          nbytes, _stream(hidden)), "splade_head_bwd_dhidden")
    return d_hidden, bitmap


def _grad_buffers(accumulate_into, V, H, dev):
    if accumulate_into is None:
        return (torch.empty((V, H), dtype=torch.float32, device=dev),
                torch.empty((V,), dtype=torch.float32, device=dev))
    dW, dbias = accumulate_into
    assert dW.dtype == torch.float32 and dW.is_contiguous() and tuple(dW.shape) == (V, H)
    assert dbias.dtype == torch.float32 and dbias.is_contiguous() and tuple(dbias.shape) == (V,)
    return dW, dbias


def head_backward_dw(segments, H: int, V: int, accumulate_into=None):
    """Phase 2: ONE dW / dbias sweep over ``segments`` = [(g_rep, rep, argmax, hidden, bitmap), ...]
    (``splade_head_bwd_dw``).  ``accumulate_into=(dW, dbias)`` adds into existing fp32 buffers."""
    lib = load_library()
    n = len(segments)
    assert 1 <= n <= _lib.HEAD_BWD_MAX_SEGMENTS
    dev = segments[0][3].device
    with torch.cuda.device(dev):
        dW, dbias = _grad_buffers(accumulate_into, V, H, dev)
        segs = (_lib.HeadBwdSegment * n)()
        maps = (ctypes.c_void_p * n)()
        for i, (g_rep, rep, argmax, hidden, bitmap) in enumerate(segments):
            segs[i] = _segment_struct(g_rep, rep, argmax, hidden, None)
            maps[i] = bitmap.data_ptr()
        check(lib.splade_head_bwd_dw(segs, maps, n, H, V, dW.data_ptr(), dbias.data_ptr(),
                                     int(accumulate_into is not None), _stream(segments[0][3])), "splade_head_bwd_dw")
    return dW, dbias


def head_backward_last(last, segments, weight: torch.Tensor, want_d_hidden: bool = True, accumulate_into=None):
    """What autograd's last head node of a backward pass runs (``splade_head_bwd_last``): dHidden of ``last`` =
    (g_rep, rep, argmax, hidden, bitmap) and the dW / dbias sweep over ``segments`` (same 5-tuples, ``last`` among
    them), the sweep concurrently on a side stream.  -> (d_hidden | None, dW, dbias)."""
    lib = load_library()
    n = len(segments)
    assert 1 <= n <= _lib.HEAD_BWD_MAX_SEGMENTS
    g_rep, rep, argmax, hidden, bitmap = last
    B, S, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    with torch.cuda.device(dev):
        d_hidden = torch.empty((B, S, H), dtype=torch.float32, device=dev) if want_d_hidden else None
        dW, dbias = _grad_buffers(accumulate_into, V, H, dev)
        segs = (_lib.HeadBwdSegment * n)()
        maps = (ctypes.c_void_p * n)()
        for i, (g_i, rep_i, arg_i, hid_i, bm_i) in enumerate(segments):
            segs[i] = _segment_struct(g_i, rep_i, arg_i, hid_i, None)
            maps[i] = bm_i.data_ptr()
        seg_last = _segment_struct(g_rep, rep, argmax, hidden, d_hidden, bitmap)
        nbytes = lib.splade_workspace_bytes(_lib.OP_HEAD_BWD_DHIDDEN, B, S, H, V, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_head_bwd_last(ctypes.byref(seg_last), segs, maps, n, weight.data_ptr(), H, V, None,
                                       dW.data_ptr(), dbias.data_ptr(), int(accumulate_into is not None), ws_ptr,
                                       nbytes, _stream(hidden)), "splade_head_bwd_last")
    return d_hidden, dW, dbias


def head_backward_multi(segments, weight: torch.Tensor, accumulate_into=None):
    """All phases for several head calls in three launches (``splade_head_bwd_multi``).
    ``segments`` = [(g_rep, rep, argmax, hidden[, active bitmap]), ...] -> ([d_hidden_i], dW, dbias)."""
    lib = load_library()
    n = len(segments)
    assert 1 <= n <= _lib.HEAD_BWD_MAX_SEGMENTS
    V, H = weight.shape
    dev = weight.device
    with torch.cuda.device(dev):
        dW, dbias = _grad_buffers(accumulate_into, V, H, dev)
        segs = (_lib.HeadBwdSegment * n)()
        keep, d_hiddens = [], []
        for i, seg in enumerate(segments):
            g_rep, rep, argmax, hidden = seg[:4]
            g_rep = _f32c(g_rep)
            keep.append(g_rep)
            dh = torch.empty(tuple(hidden.shape), dtype=torch.float32, device=dev)
            d_hiddens.append(dh)
            segs[i] = _segment_struct(g_rep, rep, argmax, hidden, dh, seg[4] if len(seg) > 4 else None)
        nbytes = lib.splade_head_bwd_multi_workspace_bytes(segs, n, V)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_head_bwd_multi(segs, n, weight.data_ptr(), H, V, dW.data_ptr(), dbias.data_ptr(),
                                        int(accumulate_into is not None), ws_ptr, nbytes, _stream(weight)),
              "splade_head_bwd_multi")
    return d_hiddens, dW, dbias


def flops_nnz(rep: torch.Tensor, out2: Optional[torch.Tensor] = None):
    """rep [N,V] fp32 -> (colmean [V], out2 [2] = {sum_j mean_j^2, mean_i nnz_i})"""
    lib = load_library()
    _require_cuda(rep, "rep")
    rep = _f32c(rep)
    N, V = rep.shape
    dev = rep.device
    with torch.cuda.device(dev):
        colmean = torch.empty((V,), dtype=torch.float32, device=dev)
        if out2 is None:
            out2 = torch.empty((2,), dtype=torch.float32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_FLOPS_NNZ, N, 0, 0, V, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_flops_nnz(rep.data_ptr(), N, V, colmean.data_ptr(), out2.data_ptr(), ws_ptr, nbytes,
                                   _stream(rep)), "splade_flops_nnz")
    return colmean, out2


def scores(q: torch.Tensor, docs: torch.Tensor, round_bf16: bool) -> torch.Tensor:
    """raw[i,j] = <q_i, docs_j>  ([Bq,V] x [N,V] -> [Bq,N] fp32)"""
    lib = load_library()
    _require_cuda(q, "q")
    q, docs = _f32c(q), _f32c(docs)
    Bq, V = q.shape
    N = docs.shape[0]
    dev = q.device
    with torch.cuda.device(dev):
        raw = torch.empty((Bq, N), dtype=torch.float32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_SCORES, Bq, 0, 0, V, N)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_scores(q.data_ptr(), Bq, docs.data_ptr(), N, V, int(round_bf16), raw.data_ptr(), ws_ptr,
                                nbytes, _stream(q)), "splade_scores")
    return raw


def rowdot(q: torch.Tensor, docs: torch.Tensor) -> torch.Tensor:
    """out[i,j] = <q_i, docs_{i,j}>  ([Bq,V] x [Bq,k,V] -> [Bq,k] fp32)"""
    lib = load_library()
    _require_cuda(q, "q")
    q, docs = _f32c(q), _f32c(docs)
    Bq, V = q.shape
    k = docs.shape[1]
    with torch.cuda.device(q.device):
        out = torch.empty((Bq, k), dtype=torch.float32, device=q.device)
        check(lib.splade_rowdot(q.data_ptr(), docs.data_ptr(), Bq, k, V, out.data_ptr(), _stream(q)),
              "splade_rowdot")
    return out


def topk(rep: torch.Tensor, k: int, blocklist: Optional[torch.Tensor] = None):
    """rep [B,V] fp32 -> (ids [B,k] int32 (-1 pad), vals [B,k] fp32, counts [B] int32); order = (weight desc,
    id asc); keeps weight > 0 and ids not in the blocklist (uint8 [V], nonzero = drop)."""
    lib = load_library()
    _require_cuda(rep, "rep")
    rep = _f32c(rep)
    B, V = rep.shape
    dev = rep.device
    if blocklist is not None:
        assert blocklist.dtype == torch.uint8 and blocklist.numel() == V and blocklist.device == dev
        blocklist = blocklist.contiguous()
    with torch.cuda.device(dev):
        ids = torch.empty((B, k), dtype=torch.int32, device=dev)
        vals = torch.empty((B, k), dtype=torch.float32, device=dev)
        counts = torch.empty((B,), dtype=torch.int32, device=dev)
        check(lib.splade_topk(rep.data_ptr(), B, V, _ptr(blocklist), k, ids.data_ptr(), vals.data_ptr(),
                              counts.data_ptr(), _stream(rep)), "splade_topk")
    return ids, vals, counts


def head_prologue_forward(x16: torch.Tensor, dense_w16: torch.Tensor, norm_w: torch.Tensor, eps: float, T: int):
    """ModernBertPredictionHead under autocast(bf16) (``splade_head_prologue_fwd``):
    x16 [rows_cap,H] bf16 (T valid rows, rows_cap >= max(T,128)) -> (y, hidden) bf16 [rows_cap,H], stats [T,2] fp32."""
    lib = load_library()
    _require_cuda(x16, "x")
    rows_cap, H = x16.shape
    assert x16.dtype == torch.bfloat16 and x16.is_contiguous() and rows_cap >= max(T, 128)
    assert dense_w16.dtype == torch.bfloat16 and dense_w16.is_contiguous() and tuple(dense_w16.shape) == (H, H)
    assert norm_w.dtype == torch.float32 and norm_w.is_contiguous() and norm_w.numel() == H
    dev = x16.device
    with torch.cuda.device(dev):
        y = torch.empty((rows_cap, H), dtype=torch.bfloat16, device=dev)
        hidden = torch.empty((rows_cap, H), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((T, 2), dtype=torch.float32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_HEAD_PROLOGUE, T, 0, H, 0, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_head_prologue_fwd(x16.data_ptr(), dense_w16.data_ptr(), norm_w.data_ptr(), float(eps), T,
                                           rows_cap, H, y.data_ptr(), hidden.data_ptr(), stats.data_ptr(), ws_ptr,
                                           nbytes, _stream(x16)), "splade_head_prologue_fwd")
    return y, hidden, stats


def head_prologue_backward(d_hidden: torch.Tensor, y: torch.Tensor, stats: torch.Tensor, norm_w: torch.Tensor,
                           dense_w16_t: Optional[torch.Tensor], T: int):
    """``splade_head_prologue_bwd``: d_hidden [T,H] fp32 -> (dy bf16 [rows_cap,H], dx bf16 [rows_cap,H] | None
    (when ``dense_w16_t`` is None), d_norm_w [H] fp32)."""
    lib = load_library()
    rows_cap, H = y.shape
    dev = y.device
    assert d_hidden.dtype == torch.float32 and d_hidden.is_contiguous() and d_hidden.numel() == T * H
    with torch.cuda.device(dev):
        dy = torch.empty((rows_cap, H), dtype=torch.bfloat16, device=dev)
        dx = torch.empty((rows_cap, H), dtype=torch.bfloat16, device=dev) if dense_w16_t is not None else None
        d_norm_w = torch.empty((H,), dtype=torch.float32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_HEAD_PROLOGUE, T, 0, H, 0, 0)
        ws, ws_ptr = _ws(nbytes, dev)
        check(lib.splade_head_prologue_bwd(d_hidden.data_ptr(), y.data_ptr(), stats.data_ptr(), norm_w.data_ptr(),
                                           _ptr(dense_w16_t), T, rows_cap, H, dy.data_ptr(), _ptr(dx),
                                           d_norm_w.data_ptr(), ws_ptr, nbytes, _stream(y)), "splade_head_prologue_bwd")
    return dy, dx, d_norm_w


# ------------------------------------------------------------------------------------------------
# autograd: fused head
# ------------------------------------------------------------------------------------------------
_WEIGHT_CACHE = []  # at most one entry: (weakref to the parameter, _version, data_ptr, bf16 copy)


def clear_weight_cache() -> None:
    """Drop the cached bf16 copy (call after mutating a weight through ``.data``, which bumps no version)."""
    _WEIGHT_CACHE.clear()
    _DENSE_CACHE.clear()


def cached_bf16(weight: torch.Tensor) -> torch.Tensor:
    """bf16 copy of a parameter, reused until the parameter is modified in place (optimizer step).
    Mirrors autocast's own weight-cast cache; saves two 77 MB casts per micro-step.

    The entry is tied to the parameter OBJECT (weak reference), not to its address: a second model whose tied
    matrix lands on the allocator address a freed model used, with the same version counter, must never see the
    first model's weights."""
    if weight.dtype == torch.bfloat16 and weight.is_contiguous():
        return weight.detach()
    if _WEIGHT_CACHE:
        ref, version, ptr, w16 = _WEIGHT_CACHE[0]
        if (ref() is weight and version == weight._version and ptr == weight.data_ptr()
                and w16.shape == weight.shape and w16.device == weight.device):
            return w16
    w = weight.detach().to(torch.bfloat16).contiguous()
    _WEIGHT_CACHE.clear()  # one live entry is enough (the tied decoder/embedding matrix)
    try:
        _WEIGHT_CACHE.append((weakref.ref(weight), weight._version, weight.data_ptr(), w))
    except TypeError:  # not weak-referenceable: do not cache
        pass
    return w


_DENSE_CACHE = {}  # id(parameter) -> (weakref, _version, data_ptr, bf16 copy, bf16 transposed copy)


def cached_dense_bf16(weight: torch.Tensor):
    """(bf16 copy, bf16 transposed copy) of ``head.dense.weight`` [H,H], reused until the parameter changes; tied to
    the parameter object like ``cached_bf16``.  The transposed copy is the K-major operand of the backward's dgrad."""
    key = id(weight)
    hit = _DENSE_CACHE.get(key)
    if hit is not None:
        ref, version, ptr, w16, w16_t = hit
        if ref() is weight and version == weight._version and ptr == weight.data_ptr():
            return w16, w16_t
    w16 = weight.detach().to(torch.bfloat16).contiguous()
    w16_t = w16.t().contiguous()
    try:
        _DENSE_CACHE[key] = (weakref.ref(weight, lambda _r, k=key: _DENSE_CACHE.pop(k, None)), weight._version,
                             weight.data_ptr(), w16, w16_t)
    except TypeError:
        pass
    return w16, w16_t


class _HeadTicket:
    """One head call whose backward is outstanding (see ``_DeferredDW``)."""

    __slots__ = ("seq", "node", "task", "segment", "__weakref__")

    def __init__(self, seq, node):
        self.seq = seq           # forward call order: the fixed summation order of the dW sweep
        self.node = node         # weakref to the autograd node (ctx) of that call
        self.task = -1           # id of the backward pass (graph task) in which its g_rep arrived
        self.segment = None      # (g_rep, rep, argmax, hidden, bitmap) once arrived


class _DeferredDW:
    """One backward, one dW.  The reference trainer calls the model three times per micro-step and back-propagates
    once (train_v33_ddp.py:339-343, 364); all three head calls share the tied decoder matrix.  Instead of three
    dense [V,H] gradients summed by autograd, every head backward computes its dHidden at once, parks its segment
    here and returns NO weight gradient, until the last head node of the running backward pass arrives: that one
    runs the single dW / dbias sweep over all parked segments (``splade_head_bwd_dw``) and returns the result.
    "Last" is decided by asking the autograd engine which of the other outstanding head nodes it will still
    execute in this pass (``torch._C._will_engine_execute_node``): nodes of an unrelated or abandoned graph
    (an evaluation forward under grad mode, a recomputation) are ignored, never waited for."""

    def __init__(self):
        self._groups = {}  # id(weight parameter) -> (weakref to it, [tickets]); identity-keyed (tensors define ==)
        self._seq = 0

    def _group(self, weight, create: bool):
        key = id(weight)
        hit = self._groups.get(key)
        if hit is not None and hit[0]() is weight:
            return hit[1]
        if not create:
            return None
        tickets = []
        self._groups[key] = (weakref.ref(weight, lambda _r, k=key: self._groups.pop(k, None)), tickets)
        return tickets

    def open(self, weight, ctx) -> _HeadTicket:
        group = self._group(weight, create=True)
        group[:] = [t for t in group if t.node() is not None]
        self._seq += 1
        t = _HeadTicket(self._seq, weakref.ref(ctx))
        group.append(t)
        return t

    def arrive(self, weight, ticket, segment):
        """-> list of segments to sweep now (forward call order), or None when another head node of this pass
        is still to come."""
        task = torch._C._current_graph_task_id()
        ticket.task, ticket.segment = task, segment
        group = self._group(weight, create=False)
        if group is None or task < 0:  # not inside the engine (direct call): nothing to wait for
            ticket.segment = None
            return [segment]
        ready = []
        for t in group:
            if t is ticket:
                ready.append(t)
                continue
            node = t.node()
            if node is None:
                continue
            if t.segment is not None and t.task == task:
                ready.append(t)       # arrived earlier in this pass
            elif t.segment is not None:
                t.segment = None      # left over from a pass that never finished (exception): drop
            elif torch._C._will_engine_execute_node(node):
                return None           # still to come in this pass
        ready.sort(key=lambda t: t.seq)
        segments = [t.segment for t in ready]
        for t in ready:
            t.segment = None
        return segments


_DEFERRED_DW = _DeferredDW()


def _defer_dw_enabled() -> bool:
    return os.environ.get("SPLADE_B200_DEFER_DW", "1") != "0"


class SpladeHeadFunction(torch.autograd.Function):
    """sparse_repr, token_weights = SPLADE-max head(hidden, decoder.weight, decoder.bias, attention_mask).

    Replaces src/model/splade_modern.py:73-86 + the decoder Linear.  Backward is the gather-driven sparse
    backward: dHidden at once (``splade_head_bwd_dhidden``), dW / dbias once per backward pass for all head calls
    that share the weight (``_DeferredDW``)."""

    @staticmethod
    def forward(ctx, hidden, weight, bias, h16, w16, b32, m32, want_token_weights: bool):
        rep, argmax, tokw = head_forward(h16, w16, b32, m32, want_token_weights)
        active = rep_bitmap(rep) if any(ctx.needs_input_grad[:3]) else rep.new_empty((0,), dtype=torch.int32)
        ctx.save_for_backward(rep, argmax, h16, w16, active)
        ctx.in_dtypes = (hidden.dtype, weight.dtype, None if bias is None else bias.dtype)
        ctx.weight_ref = None
        ctx.ticket = None
        needs_w = ctx.needs_input_grad[1] or (bias is not None and ctx.needs_input_grad[2])
        if needs_w and _defer_dw_enabled():
            ctx.weight_ref = weakref.ref(weight)
            ctx.ticket = _DEFERRED_DW.open(weight, ctx)
        if tokw is None:
            tokw = torch.zeros((0,), dtype=torch.float32, device=hidden.device)
        ctx.mark_non_differentiable(tokw)
        return rep, tokw

    @staticmethod
    def backward(ctx, g_rep, _g_tokw):
        rep, argmax, h16, w16, active = ctx.saved_tensors
        hd, wd, bd = ctx.in_dtypes
        d_hidden, dW, dbias = _head_backward_core(ctx, g_rep, rep, argmax, h16, w16, active, ctx.needs_input_grad[0],
                                                  ctx.needs_input_grad[1], bd is not None and ctx.needs_input_grad[2])
        return (d_hidden.to(hd) if d_hidden is not None else None, None if dW is None else dW.to(wd),
                None if dbias is None else dbias.to(bd), None, None, None, None, None)


def _head_backward_core(ctx, g_rep, rep, argmax, h16, w16, active, need_h: bool, need_dw: bool, need_db: bool):
    """Shared by the head Functions: -> (d_hidden fp32 | None, dW fp32 | None, dbias fp32 | None).  dHidden of this
    call at once; dW / dbias ONCE per backward pass, by the last head node of the pass (``_DeferredDW``)."""
    need_w = need_dw or need_db
    g32 = _f32c(g_rep)
    act = active if active.numel() else rep_bitmap(rep)
    dW = dbias = d_hidden = None
    segments = None
    if need_w:
        segment = (g32, rep, argmax, h16, act)
        weight = ctx.weight_ref() if ctx.weight_ref is not None else None
        segments = [segment]
        if ctx.ticket is not None and weight is not None:
            segments = _DEFERRED_DW.arrive(weight, ctx.ticket, segment)
    if segments is None:
        # not the last head node of this pass: its dHidden now, its share of dW with the others later
        d_hidden, _ = head_backward_dhidden(g32, rep, argmax, h16, w16, want_d_hidden=need_h, active=act)
    else:
        # the last head node (or the only one): bucket + dHidden of this segment and, concurrently on a side
        # stream, the ONE dW / dbias sweep over every segment of the pass
        d_hidden, dW, dbias = head_backward_last((g32, rep, argmax, h16, act), segments, w16, want_d_hidden=need_h)
        dW = dW if need_dw else None
        dbias = dbias if need_db else None
    return d_hidden, dW, dbias


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 library GEMM with an fp32 result (cuBLAS accumulates in fp32 either way)."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.mm(a, b).to(torch.float32)


class SpladeHeadWithPrologueFunction(torch.autograd.Function):
    """``SpladeHeadFunction`` preceded by the trunk's prediction head, ``hidden = LayerNorm(GELU(dense(x)))``
    (transformers modeling_modernbert.py:493-502, called at src/model/splade_modern.py:69-73), as autocast(bf16)
    computes it: one tcgen05 GEMM + one fused GELU/LayerNorm pass emitting the bf16 head input directly (no
    separate cast), and a fused LayerNorm/GELU backward + tcgen05 dgrad.  dHidden stays fp32 between the two
    backward stages (the reference rounds it to bf16 there).  The dense weight gradient is one library GEMM."""

    @staticmethod
    def forward(ctx, x, dense_w, norm_w, weight, bias, w16, b32, m32, eps: float, want_token_weights: bool):
        B, S, H = x.shape
        T = B * S
        rows_cap = max(T, 128)
        x16 = x.detach().reshape(T, H).to(torch.bfloat16)
        if rows_cap != T:
            pad = torch.zeros((rows_cap, H), dtype=torch.bfloat16, device=x.device)
            pad[:T] = x16
            x16 = pad
        x16 = x16.contiguous()
        dw16, dw16_t = cached_dense_bf16(dense_w)
        gamma = _f32c(norm_w)
        y, hidden, stats = head_prologue_forward(x16, dw16, gamma, eps, T)
        h16 = hidden[:T].view(B, S, H)
        rep, argmax, tokw = head_forward(h16, w16, b32, m32, want_token_weights)
        any_grad = any(ctx.needs_input_grad[:5])
        active = rep_bitmap(rep) if any_grad else rep.new_empty((0,), dtype=torch.int32)
        ctx.save_for_backward(rep, argmax, h16, w16, active, x16, y, stats, gamma, dw16_t)
        ctx.in_dtypes = (x.dtype, dense_w.dtype, norm_w.dtype, weight.dtype, None if bias is None else bias.dtype)
        ctx.shape = (B, S, H, T)
        ctx.weight_ref = None
        ctx.ticket = None
        needs_w = ctx.needs_input_grad[3] or (bias is not None and ctx.needs_input_grad[4])
        if needs_w and _defer_dw_enabled():
            ctx.weight_ref = weakref.ref(weight)
            ctx.ticket = _DEFERRED_DW.open(weight, ctx)
        if tokw is None:
            tokw = torch.zeros((0,), dtype=torch.float32, device=x.device)
        ctx.mark_non_differentiable(tokw, h16)
        return rep, tokw, h16

    @staticmethod
    def backward(ctx, g_rep, _g_tokw, _g_h16):
        rep, argmax, h16, w16, active, x16, y, stats, gamma, dw16_t = ctx.saved_tensors
        xd, dd, nd, wd, bd = ctx.in_dtypes
        B, S, H, T = ctx.shape
        need_x, need_dense, need_norm = ctx.needs_input_grad[:3]
        need_pro = need_x or need_dense or need_norm
        d_hidden, dW, dbias = _head_backward_core(ctx, g_rep, rep, argmax, h16, w16, active, need_pro,
                                                  ctx.needs_input_grad[3],
                                                  bd is not None and ctx.needs_input_grad[4])
        dx = d_dense = d_norm = None
        if need_pro:
            dy, dx16, d_norm = head_prologue_backward(d_hidden.view(T, H), y, stats, gamma,
                                                      dw16_t if need_x else None, T)
            if need_x:
                dx = dx16[:T].view(B, S, H).to(xd)
            if need_dense:  # dWd[o,i] = sum_t dy[t,o] x[t,i]: a plain [H,T]x[T,H] library GEMM, fp32 out
                d_dense = _mm_f32(dy[:T].t(), x16[:T]).to(dd)
            d_norm = d_norm.to(nd) if need_norm else None
        return (dx, d_dense, d_norm, None if dW is None else dW.to(wd), None if dbias is None else dbias.to(bd),
                None, None, None, None, None)


def autocast_is_bf16() -> bool:
    """True inside ``torch.autocast("cuda", dtype=torch.bfloat16)`` - the only autocast mode this path mirrors
    (src/train/cli/train_v33_ddp.py:337).  fp16 autocast is refused rather than silently computed as bf16."""
    if not torch.is_autocast_enabled("cuda"):
        return False
    dt = torch.get_autocast_dtype("cuda")
    if dt != torch.bfloat16:
        raise RuntimeError(f"splade_b200: autocast dtype {dt} is not supported; the fused path mirrors "
                           f"torch.autocast('cuda', dtype=torch.bfloat16) only")
    return True


class LazyTokenWeights(torch.Tensor):
    """``token_weights [B,S]`` that is computed only if somebody reads it.

    Every caller in the reference discards the second return value of ``SPLADEModernBERT.forward``
    (``anchor_repr, _ = model(...)``, src/train/cli/train_v33_ddp.py:339-343; ``encode`` drops it at
    splade_modern.py:96), yet ``forward`` must return it, and the per-position reduction over the vocabulary costs
    11 % of the fused kernel (30 % on query batches).  This wrapper tensor reports shape / dtype / device without
    touching the GPU; the first operation on it runs the forward kernel variant that also produces token_weights
    (bit-identical to the eager path) and every operation is then forwarded to the real tensor."""

    @staticmethod
    def __new__(cls, shape, device, thunk):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=torch.float32, device=device,
                                                requires_grad=False)
        t._thunk, t._value = thunk, None
        return t

    def materialize(self) -> torch.Tensor:
        if self._value is None:
            self._value = self._thunk()
            self._thunk = None  # drop the references to the forward's operands
        return self._value

    @property
    def is_materialized(self) -> bool:
        return self._value is not None

    def __repr__(self):
        return (f"LazyTokenWeights(shape={tuple(self.shape)}, device={self.device}, "
                f"materialized={self._value is not None})")

    # accessors implemented in C against the storage (a wrapper subclass has none)
    def numpy(self, *a, **k):
        return self.materialize().numpy(*a, **k)

    def tolist(self):
        return self.materialize().tolist()

    def item(self):
        return self.materialize().item()

    def data_ptr(self):
        return self.materialize().data_ptr()

    def __array__(self, *a, **k):
        return self.materialize().__array__(*a, **k)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        from torch.utils._pytree import tree_map

        def unwrap(x):
            return x.materialize() if isinstance(x, LazyTokenWeights) else x

        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))


_WARNED_FP32 = False


def _warn_fp32_once() -> None:
    """Outside autocast the reference's decoder Linear runs in fp32 (benchmark/encoders.py:309-318); the fused head
    always multiplies bf16 operands (fp32 accumulate, logits rounded once to bf16).  Say so once."""
    global _WARNED_FP32
    if not _WARNED_FP32:
        _WARNED_FP32 = True
        import warnings
        warnings.warn("splade_b200: the fused SPLADE head computes with bf16 operands also outside torch.autocast "
                      "(the reference is fp32 there): weights agree within atol 1e-2 + rtol 1e-2, top-64 term sets "
                      "overlap >= 95 %; wrap the call in torch.autocast('cuda', dtype=torch.bfloat16) to mirror the "
                      "reference's own bf16 recipe bit for bit.", stacklevel=3)


def lazy_token_weights_enabled() -> bool:
    return os.environ.get("SPLADE_B200_LAZY_TOKEN_WEIGHTS", "1") != "0"


def head_prologue_enabled() -> bool:
    return os.environ.get("SPLADE_B200_HEAD_PROLOGUE", "1") != "0"


def prologue_of(head_module) -> Optional[tuple]:
    """(dense.weight, norm.weight, eps) of a transformers ``ModernBertPredictionHead`` when the fused prologue can
    stand in for it - no biases (classifier_bias = norm_bias = false), exact-erf GELU, hidden size 256..1024 in
    steps of 256 - else None (the module itself runs, as before)."""
    dense, norm, act = getattr(head_module, "dense", None), getattr(head_module, "norm", None), \
        getattr(head_module, "act", None)
    if not isinstance(dense, torch.nn.Linear) or not isinstance(norm, torch.nn.LayerNorm):
        return None
    if dense.bias is not None or norm.bias is not None or norm.weight is None:
        return None
    H = dense.in_features
    if dense.out_features != H or H % 256 != 0 or not 256 <= H <= 1024 or tuple(norm.normalized_shape) != (H,):
        return None
    cfg_act = getattr(getattr(head_module, "config", None), "classifier_activation", None)
    exact_gelu = cfg_act == "gelu" or (isinstance(act, torch.nn.GELU) and act.approximate == "none")
    if not exact_gelu:
        return None
    return dense.weight, norm.weight, float(norm.eps)


def splade_head(hidden, weight, bias, attention_mask, want_token_weights: bool = True,
                round_bias: Optional[bool] = None, lazy_token_weights: Optional[bool] = None,
                prologue: Optional[tuple] = None):
    """(sparse_repr [B,V] fp32, token_weights [B,S] fp32 | None).  ``lazy_token_weights`` (default: env
    SPLADE_B200_LAZY_TOKEN_WEIGHTS, on): token_weights is a ``LazyTokenWeights`` - computed on first use.
    ``prologue`` = ``prologue_of(model.head)``: ``hidden`` is then the TRUNK output and the prediction head
    (dense -> GELU -> LayerNorm) runs inside this call (bf16 autocast arithmetic; ``SpladeHeadWithPrologueFunction``)."""
    _require_cuda(hidden, "hidden")
    if round_bias is None:
        round_bias = autocast_is_bf16()
        if not round_bias and hidden.dtype == torch.float32:
            _warn_fp32_once()
    if lazy_token_weights is None:
        lazy_token_weights = lazy_token_weights_enabled()
    w16 = cached_bf16(weight)
    if bias is None:
        b32 = torch.zeros(weight.shape[0], dtype=torch.float32, device=hidden.device)
    else:
        b32 = bias.detach()
        if round_bias:  # autocast hands F.linear a bf16 bias
            b32 = b32.to(torch.bfloat16)
        b32 = b32.to(torch.float32).contiguous()
    m32 = attention_mask.detach().to(torch.int32).contiguous()
    eager = want_token_weights and not lazy_token_weights
    if prologue is not None:
        dense_w, norm_w, eps = prologue
        rep, tokw, h16 = SpladeHeadWithPrologueFunction.apply(hidden, dense_w, norm_w, weight, bias, w16, b32, m32,
                                                              eps, eager)
    else:
        h16 = hidden.detach().to(torch.bfloat16).contiguous()
        rep, tokw = SpladeHeadFunction.apply(hidden, weight, bias, h16, w16, b32, m32, eager)
    if not want_token_weights:
        return rep, None
    if eager:
        return rep, tokw
    B, S, _ = h16.shape
    return rep, LazyTokenWeights((B, S), hidden.device, lambda: head_forward(h16, w16, b32, m32, True)[2])


# ------------------------------------------------------------------------------------------------
# autograd: SPLADELossV33 on the pooled vectors
# ------------------------------------------------------------------------------------------------
def _loss_bwd_tc_wanted(n_docs: int) -> bool:
    """Autocast mode: run the backward's two products on tcgen05 (splade_loss_bwd_tc)?  On by default
    (profiles/r02_time_loss.log: 43 vs 59 us at N = 64 in-batch documents, 82 vs 236 us at N = 512);
    SPLADE_B200_LOSS_BWD_TC=0 keeps the fp32 CUDA-core kernels."""
    return os.environ.get("SPLADE_B200_LOSS_BWD_TC", "1") != "0"


class V33LossFunction(torch.autograd.Function):
    """total loss + packed scalars from (anchor [Bq,V], docs [N,V], negatives [Bq,k,V]).

    Forward = src/model/losses.py:209-269; ``scalars`` (non-differentiable, one D2H copy on the caller's side)
    holds {loss, infonce, flops_q, flops_d, flops_neg, kd, margin_mse, nonzero_q, nonzero_d}."""

    @staticmethod
    def forward(ctx, anchor, docs, negatives, teacher_kd, teacher_pos, teacher_neg, cfg: dict):
        lib = load_library()
        _require_cuda(anchor, "anchor_repr")
        dev = anchor.device
        q = _f32c(anchor)
        d = _f32c(docs)
        n = _f32c(negatives)
        Bq, V = q.shape
        N = d.shape[0]
        k = n.shape[1]
        off = int(cfg["label_offset"])
        round_bf16 = bool(cfg["round_bf16"])
        has_kd = teacher_kd is not None and cfg["lambda_kd"] > 0
        has_mmse = teacher_pos is not None and teacher_neg is not None and cfg["lambda_mmse"] > 0
        with torch.cuda.device(dev):
            colmean = torch.empty((3, V), dtype=torch.float32, device=dev)
            scalars = torch.empty((16,), dtype=torch.float32, device=dev)
            g_in = torch.empty((Bq, N), dtype=torch.float32, device=dev)
            g_hard = torch.empty((Bq, k), dtype=torch.float32, device=dev)
            t_kd = _f32c(teacher_kd) if has_kd else None
            t_pos = _f32c(teacher_pos).view(Bq) if has_mmse else None
            t_neg = _f32c(teacher_neg).view(Bq, k) if has_mmse else None
            nbytes = lib.splade_workspace_bytes(_lib.OP_V33_LOSS_FWD, Bq, N, 0, V, k)
            ws, ws_ptr = _ws(nbytes, dev)
            args = _lib.V33FwdArgs(
                Bq=Bq, N=N, k=k, V=V, label_offset=off, round_bf16=int(round_bf16), has_kd=int(has_kd),
                has_mmse=int(has_mmse),
                temperature=cfg["temperature"], kd_temperature=cfg["kd_temperature"],
                lambda_q=cfg["lambda_q"], lambda_d=cfg["lambda_d"], lambda_neg=cfg["lambda_neg"],
                lambda_kd=cfg["lambda_kd"], lambda_mmse=cfg["lambda_mmse"], reserved_f=0.0,
                q=q.data_ptr(), docs=d.data_ptr(), negs=n.data_ptr(),
                teacher_kd=_ptr(t_kd), teacher_pos=_ptr(t_pos), teacher_neg=_ptr(t_neg),
                colmean=colmean.data_ptr(), scalars=scalars.data_ptr(), g_inbatch=g_in.data_ptr(),
                g_hard=g_hard.data_ptr(), workspace=ws_ptr, ws_bytes=nbytes)
            check(lib.splade_v33_loss_fwd(ctypes.byref(args), _stream(q)), "splade_v33_loss_fwd")
            cm_q, cm_d, cm_n = colmean[0], colmean[1], colmean[2]
            # autocast mode: the forward leaves the bf16 operands of the score GEMM in its workspace; the backward's
            # two products run on them with tcgen05 (splade_loss_bwd_tc), so the workspace stays alive until then
            ctx.tc = None
            if round_bf16 and V % 4 == 0 and _loss_bwd_tc_wanted(N):
                q16, d16, vp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int(0)
                check(lib.splade_v33_loss_fwd_operands(ctypes.byref(args), ctypes.byref(q16), ctypes.byref(d16),
                                                       ctypes.byref(vp)), "splade_v33_loss_fwd_operands")
                ctx.tc = (ws, q16.value, d16.value)
        ctx.save_for_backward(q, d, n, g_in, g_hard, cm_q, cm_d, cm_n)
        ctx.meta = (Bq, N, k, V, off, cfg["lambda_q"], cfg["lambda_d"], cfg["lambda_neg"],
                    anchor.dtype, docs.dtype, negatives.dtype, tuple(negatives.shape))
        loss = scalars[0].clone()
        ctx.mark_non_differentiable(scalars)
        return loss, scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        lib = load_library()
        q, d, n, g_in, g_hard, cm_q, cm_d, cm_n = ctx.saved_tensors
        Bq, N, k, V, off, lq, ld, ln, qd, dd, nd, nshape = ctx.meta
        dev = q.device
        gl = _f32c(g_loss).reshape(1)
        with torch.cuda.device(dev):
            dq = torch.empty_like(q)
            ddocs = torch.empty_like(d)
            dnegs = torch.empty_like(n)
            if ctx.tc is not None:
                _keep, q16, d16 = ctx.tc
                nbytes = lib.splade_workspace_bytes(_lib.OP_LOSS_BWD_TC, Bq, 0, 0, 0, N)
                ws, ws_ptr = _ws(nbytes, dev)
                check(lib.splade_loss_bwd_tc(q.data_ptr(), n.data_ptr(), q16, d16, Bq, N, k, V, off,
                                             g_in.data_ptr(), g_hard.data_ptr(), cm_q.data_ptr(), cm_d.data_ptr(),
                                             cm_n.data_ptr(), 2.0 * lq / Bq, 2.0 * ld / Bq, 2.0 * ln / (Bq * k),
                                             gl.data_ptr(), dq.data_ptr(), ddocs.data_ptr(), dnegs.data_ptr(),
                                             ws_ptr, nbytes, _stream(q)), "splade_loss_bwd_tc")
            else:
                check(lib.splade_loss_bwd(q.data_ptr(), d.data_ptr(), n.data_ptr(), Bq, N, k, V, off,
                                          g_in.data_ptr(), g_hard.data_ptr(), cm_q.data_ptr(), cm_d.data_ptr(),
                                          cm_n.data_ptr(), 2.0 * lq / Bq, 2.0 * ld / Bq, 2.0 * ln / (Bq * k),
                                          gl.data_ptr(), dq.data_ptr(), ddocs.data_ptr(), dnegs.data_ptr(),
                                          _stream(q)), "splade_loss_bwd")
        return dq.to(qd), ddocs.to(dd), dnegs.view(nshape).to(nd), None, None, None, None


class V33PeerLossFunction(torch.autograd.Function):
    """``V33LossFunction`` with the positives of every rank as in-batch documents, exchanged through
    ``PeerExchange`` windows by the kernels themselves (``splade_peer_v33_loss_fwd`` / ``splade_peer_loss_bwd``):
    no NCCL all-gather forward, no ``[W*B, V]`` gradient and no reduce-scatter backward.  Returns the gradient
    of the LOCAL positives summed over every rank's queries - what all-gather + reduce-scatter(sum) produce."""

    @staticmethod
    def forward(ctx, anchor, docs_local, negatives, teacher_pos, teacher_neg, cfg: dict, exchange):
        lib = load_library()
        _require_cuda(anchor, "anchor_repr")
        dev = anchor.device
        q = _f32c(anchor)
        d = _f32c(docs_local)
        n = _f32c(negatives)
        Bq, V = q.shape
        k = n.shape[1]
        W = exchange.world
        if (Bq, V) != (exchange.batch, exchange.vocab) or tuple(d.shape) != (Bq, V):
            raise RuntimeError(f"SPLADELossV33: the peer exchange was built for batch {exchange.batch} x vocabulary "
                               f"{exchange.vocab}, got queries {tuple(q.shape)} / positives {tuple(d.shape)}")
        round_bf16 = bool(cfg["round_bf16"])
        has_mmse = teacher_pos is not None and teacher_neg is not None and cfg["lambda_mmse"] > 0
        epoch = exchange.next_epoch()
        with torch.cuda.device(dev):
            colmean = torch.empty((3, V), dtype=torch.float32, device=dev)
            scalars = torch.empty((16,), dtype=torch.float32, device=dev)
            g_hard = torch.empty((Bq, k), dtype=torch.float32, device=dev)
            docs_all = torch.empty((W * Bq, V), dtype=torch.float32, device=dev)
            t_pos = _f32c(teacher_pos).view(Bq) if has_mmse else None
            t_neg = _f32c(teacher_neg).view(Bq, k) if has_mmse else None
            nbytes = lib.splade_workspace_bytes(_lib.OP_V33_LOSS_FWD, Bq, W * Bq, 0, V, k)
            ws, ws_ptr = _ws(nbytes, dev)
            args = _lib.V33FwdArgs(
                Bq=Bq, N=Bq, k=k, V=V, label_offset=0, round_bf16=int(round_bf16), has_kd=0,
                has_mmse=int(has_mmse),
                temperature=cfg["temperature"], kd_temperature=cfg["kd_temperature"],
                lambda_q=cfg["lambda_q"], lambda_d=cfg["lambda_d"], lambda_neg=cfg["lambda_neg"],
                lambda_kd=0.0, lambda_mmse=cfg["lambda_mmse"], reserved_f=0.0,
                q=q.data_ptr(), docs=d.data_ptr(), negs=n.data_ptr(),
                teacher_kd=None, teacher_pos=_ptr(t_pos), teacher_neg=_ptr(t_neg),
                colmean=colmean.data_ptr(), scalars=scalars.data_ptr(), g_inbatch=None,
                g_hard=g_hard.data_ptr(), workspace=ws_ptr, ws_bytes=nbytes)
            check(lib.splade_peer_v33_loss_fwd(ctypes.byref(args), ctypes.byref(exchange.group_struct), epoch,
                                               docs_all.data_ptr(), _stream(q)), "splade_peer_v33_loss_fwd")
            ctx.tc = None
            if round_bf16 and _loss_bwd_tc_wanted(W * Bq):
                q16, d16, vp = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int(0)
                check(lib.splade_v33_loss_fwd_operands(ctypes.byref(args), ctypes.byref(q16), ctypes.byref(d16),
                                                       ctypes.byref(vp)), "splade_v33_loss_fwd_operands")
                ctx.tc = (ws, q16.value, d16.value)  # the forward workspace (bf16 operands) lives until the backward
        ctx.save_for_backward(q, docs_all, n, g_hard, colmean)
        ctx.exchange, ctx.epoch = exchange, epoch
        ctx.meta = (Bq, k, V, cfg["lambda_q"], cfg["lambda_d"], cfg["lambda_neg"],
                    anchor.dtype, docs_local.dtype, negatives.dtype, tuple(negatives.shape))
        loss = scalars[0].clone()
        ctx.mark_non_differentiable(scalars)
        return loss, scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        lib = load_library()
        q, docs_all, n, g_hard, colmean = ctx.saved_tensors
        Bq, k, V, lq, ld, ln, qd, dd, nd, nshape = ctx.meta
        x = ctx.exchange
        dev = q.device
        gl = _f32c(g_loss).reshape(1)
        with torch.cuda.device(dev):
            dq = torch.empty_like(q)
            ddocs = torch.empty_like(q)
            dnegs = torch.empty_like(n)
            nbytes = lib.splade_workspace_bytes(_lib.OP_PEER_LOSS_BWD, Bq, 0, 0, 0, x.world)
            ws, ws_ptr = _ws(nbytes, dev)
            check(lib.splade_peer_loss_bwd(ctypes.byref(x.group_struct), ctx.epoch, q.data_ptr(),
                                           docs_all.data_ptr(), n.data_ptr(), Bq, k, V, g_hard.data_ptr(),
                                           colmean[0].data_ptr(), colmean[1].data_ptr(), colmean[2].data_ptr(),
                                           2.0 * lq / Bq, 2.0 * ld / Bq, 2.0 * ln / (Bq * k), gl.data_ptr(),
                                           dq.data_ptr(), ddocs.data_ptr(), dnegs.data_ptr(),
                                           None if ctx.tc is None else ctx.tc[1],
                                           None if ctx.tc is None else ctx.tc[2], ws_ptr, nbytes,
                                           _stream(q)), "splade_peer_loss_bwd")
        return dq.to(qd), ddocs.to(dd), dnegs.view(nshape).to(nd), None, None, None, None
