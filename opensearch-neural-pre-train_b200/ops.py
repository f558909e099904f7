"""Torch-tensor wrappers and autograd Functions over the C-ABI (``include/splade_b200.h``).

PyTorch is plumbing here: it allocates every buffer (inputs, outputs, workspaces), supplies the current CUDA
stream and records the autograd graph.  All arithmetic happens in ``libsplade_b200.so``.
"""
from __future__ import annotations

import ctypes
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


# ------------------------------------------------------------------------------------------------
# autograd: fused head
# ------------------------------------------------------------------------------------------------
_WEIGHT_CACHE = []  # at most one entry: (weakref to the parameter, _version, data_ptr, bf16 copy)


def clear_weight_cache() -> None:
    """Drop the cached bf16 copy (call after mutating a weight through ``.data``, which bumps no version)."""
    _WEIGHT_CACHE.clear()


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


class SpladeHeadFunction(torch.autograd.Function):
    """sparse_repr, token_weights = SPLADE-max head(hidden, decoder.weight, decoder.bias, attention_mask).

    Replaces src/model/splade_modern.py:73-86 + the decoder Linear.  Backward is the gather-driven sparse
    backward (splade_head_bwd)."""

    @staticmethod
    def forward(ctx, hidden, weight, bias, attention_mask, want_token_weights: bool, round_bias: bool):
        _require_cuda(hidden, "hidden")
        h16 = hidden.detach().to(torch.bfloat16).contiguous()
        w16 = cached_bf16(weight)
        if bias is None:
            b32 = torch.zeros(weight.shape[0], dtype=torch.float32, device=hidden.device)
        else:
            b32 = bias.detach()
            if round_bias:  # autocast hands F.linear a bf16 bias
                b32 = b32.to(torch.bfloat16)
            b32 = b32.to(torch.float32).contiguous()
        m32 = attention_mask.detach().to(torch.int32).contiguous()
        rep, argmax, tokw = head_forward(h16, w16, b32, m32, want_token_weights)
        ctx.save_for_backward(rep, argmax, h16, w16)
        ctx.in_dtypes = (hidden.dtype, weight.dtype, None if bias is None else bias.dtype)
        if tokw is None:
            tokw = torch.zeros((0,), dtype=torch.float32, device=hidden.device)
        ctx.mark_non_differentiable(tokw)
        return rep, tokw

    @staticmethod
    def backward(ctx, g_rep, _g_tokw):
        rep, argmax, h16, w16 = ctx.saved_tensors
        d_hidden, dW, dbias = head_backward(g_rep, rep, argmax, h16, w16)
        hd, wd, bd = ctx.in_dtypes
        return (d_hidden.to(hd), dW.to(wd), None if bd is None else dbias.to(bd), None, None, None)


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


def splade_head(hidden, weight, bias, attention_mask, want_token_weights: bool = True,
                round_bias: Optional[bool] = None):
    if round_bias is None:
        round_bias = autocast_is_bf16() if hidden.is_cuda else False
    rep, tokw = SpladeHeadFunction.apply(hidden, weight, bias, attention_mask, want_token_weights, round_bias)
    return rep, (tokw if want_token_weights else None)


# ------------------------------------------------------------------------------------------------
# autograd: SPLADELossV33 on the pooled vectors
# ------------------------------------------------------------------------------------------------
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
                                           dq.data_ptr(), ddocs.data_ptr(), dnegs.data_ptr(), ws_ptr, nbytes,
                                           _stream(q)), "splade_peer_loss_bwd")
        return dq.to(qd), ddocs.to(dd), dnegs.view(nshape).to(nd), None, None, None, None
