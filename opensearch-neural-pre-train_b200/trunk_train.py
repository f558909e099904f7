"""Native ModernBERT trunk with autograd (training side) - opt-in.

The reference trains through transformers' ``ModernBertModel`` (src/model/splade_modern.py:69-73 under
``torch.autocast(bf16)``, src/train/cli/train_v33_ddp.py:337-364).  Measured on a B200, that trunk is 165 of the 171 ms
of a training micro-step and is GPU-bound at ~10 % of the tensor peak (replaying it from CUDA graphs gains 8 %:
tools/graph_trunk_probe.py).  ``NativeTrunkTrain`` runs the same network on the packed real tokens:

* forward: ``splade_trunk_train_fwd`` - the kernels of the inference trunk (tcgen05 GEMMs with fused RoPE / residual /
  GeGLU epilogues, tcgen05 attention), one C-ABI call, writing what the backward needs into save buffers;
* backward (this file drives it, layer by layer): dgrad GEMMs on the same tcgen05 kernel with transposed bf16
  weights (``splade_trunk_gemm``), wgrad GEMMs ``dY^T X`` on cuBLAS (plain library GEMMs, fp32 results),
  ``splade_trunk_ln_bwd`` (fp32 LayerNorm backward on the residual stream, deterministic dgamma),
  ``splade_trunk_geglu_bwd``, the attention backward through the library attention (``F.scaled_dot_product_attention``
  recomputed on the saved q / k / v in the padded layout, with the same key-padding / sliding-window visibility) and
  the RoPE transpose; embedding gradients are scatter-added into the tied matrix.

Arithmetic mirrors autograd under autocast(bf16): bf16 Linear outputs and input gradients, fp32 residual-stream
gradient, fp32 LayerNorm backward.  The parameters of the wrapped module stay the single source of truth and receive
their gradients through ``torch.autograd`` (so ``DistributedDataParallel`` hooks fire as usual).
Limits: the ModernBERT configuration the inference trunk supports, hidden size 256/512/768/1024, no dropout,
right-padded (suffix) attention masks.
"""
from __future__ import annotations

import ctypes
from typing import List

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, load_library
from .ops import _mm_f32
from .trunk import NativeTrunk


def _ln_bwd(dy: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, eps: float, dx: torch.Tensor, accumulate: bool):
    """dx (+)= LayerNorm_backward(dy; x, gamma) over all rows; -> dgamma [H] fp32 (``splade_trunk_ln_bwd``)."""
    lib = load_library()
    rows, H = x.shape
    assert x.dtype == torch.float32 and dx.dtype == torch.float32 and x.is_contiguous() and dx.is_contiguous()
    assert dy.is_contiguous() and tuple(dy.shape) == (rows, H) and dy.dtype in (torch.bfloat16, torch.float32)
    dev = x.device
    dgamma = torch.empty((H,), dtype=torch.float32, device=dev)
    nbytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_LN_BWD, 0, 0, H, 0, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    check(lib.splade_trunk_ln_bwd(dy.data_ptr(), int(dy.dtype == torch.float32), x.data_ptr(), gamma.data_ptr(),
                                  float(eps), rows, H, dx.data_ptr(), int(accumulate), None, dgamma.data_ptr(),
                                  ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream),
          "splade_trunk_ln_bwd")
    return dgamma


def _geglu_bwd(h: torch.Tensor, dgg: torch.Tensor) -> torch.Tensor:
    """h = [a | g] [rows,2I] bf16, dgg [rows,I] bf16 -> dh [rows,2I] bf16 (``splade_trunk_geglu_bwd``)."""
    lib = load_library()
    rows, I2 = h.shape
    dh = torch.empty_like(h)
    check(lib.splade_trunk_geglu_bwd(h.data_ptr(), dgg.data_ptr(), rows, I2 // 2, dh.data_ptr(),
                                     torch.cuda.current_stream(h.device).cuda_stream), "splade_trunk_geglu_bwd")
    return dh


def _dqkv_pack(dq, dk, dv, row_of, cos, sin, rows: int, H: int, out=None) -> torch.Tensor:
    """(dq, dk, dv) [B,nh,S,64] bf16 (any batch / head / position strides) -> dqkv [rows,3H] bf16 with the RoPE
    rotation transposed (``splade_trunk_dqkv_pack``).  ``row_of`` None = token (b, s) is row b*S+s."""
    lib = load_library()
    B, nh, S, D = dq.shape
    assert D == 64 and dq.dtype == torch.bfloat16
    if not (dq.stride() == dk.stride() == dv.stride() and dq.stride(3) == 1
            and all(st % 8 == 0 for st in dq.stride()[:3])):
        dq, dk, dv = dq.contiguous(), dk.contiguous(), dv.contiguous()   # one common 16-byte-aligned layout
    dev = dq.device
    if out is None:
        out = (torch.empty if row_of is None else torch.zeros)((rows, 3 * H), dtype=torch.bfloat16, device=dev)
    check(lib.splade_trunk_dqkv_pack(dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dq.stride(0), dq.stride(1),
                                     dq.stride(2), None if row_of is None else row_of.data_ptr(), cos.data_ptr(),
                                     sin.data_ptr(), B, S, nh, rows, out.data_ptr(),
                                     torch.cuda.current_stream(dev).cuda_stream), "splade_trunk_dqkv_pack")
    return out


class _TrunkTrainFunction(torch.autograd.Function):
    """hidden [B,S,H] bf16 = final_norm(ModernBertModel layers(embeddings(input_ids))) with a hand-driven backward."""

    @staticmethod
    def forward(ctx, owner, input_ids, attention_mask, *params):
        lib = load_library()
        cfg = owner.cfg
        dev = input_ids.device
        c = owner.weights(dev)
        B, S = input_ids.shape
        H, I, nh, L = cfg.hidden_size, int(cfg.intermediate_size), cfg.num_attention_heads, len(c["layers"])
        ids = input_ids.to(torch.int64).contiguous()
        m32 = attention_mask.to(torch.int32).contiguous()
        rows = int(lib.splade_packed_rows(B, S))
        # a batch without padding whose rows fill the packed buffers exactly: packed row of (b, s) is b*S+s, no row is
        # ever left unwritten, so nothing has to be zero-filled and the backward needs no gathers (one sync to know)
        dense = rows == B * S and bool((m32 != 0).all())
        with torch.cuda.device(dev):
            z = lambda *shape, dtype: torch.empty(shape, dtype=dtype, device=dev)  # noqa: E731
            sv = {"x_attn_in": z(L, rows, H, dtype=torch.float32), "x_mlp_in": z(L, rows, H, dtype=torch.float32),
                  "xb": z(L, rows, H, dtype=torch.bfloat16), "qkv": z(L, rows, 3 * H, dtype=torch.bfloat16),
                  "ab": z(L, rows, H, dtype=torch.bfloat16), "xb2": z(L, rows, H, dtype=torch.bfloat16),
                  "gg": z(L, rows, I, dtype=torch.bfloat16), "x_final": z(rows, H, dtype=torch.float32),
                  "out": z(rows + 8, H, dtype=torch.bfloat16),  # row `rows` = the zero row padding positions read
                  "meta": z(rows, dtype=torch.int32), "plan": z(8, dtype=torch.int32)}
            saves = _lib.TrunkTrainSaves(**{k: v.data_ptr() for k, v in sv.items()})
            nbytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_TRAIN_FWD, B, S, H, 0, 0)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            args = _lib.TrunkArgs(
                B=B, S=S, H=H, I=I, V=cfg.vocab_size, n_heads=nh, n_layers=L,
                local_window=int(cfg.local_attention) // 2, norm_eps=float(cfg.norm_eps),
                global_theta=c["global_theta"], local_theta=c["local_theta"], reserved_f=0.0,
                input_ids=ids.data_ptr(), mask=m32.data_ptr(), tok_emb=c["tok_emb"].data_ptr(),
                emb_norm=c["emb_norm"].data_ptr(), layers=c["layer_array"], final_norm=c["final_norm"].data_ptr(),
                head_dense=None, head_norm=None, dec_w=None, dec_bias=None, rep=None, argmax=None, hidden_out=None,
                meta_out=None, plan_out=None, workspace=ws.data_ptr(), ws_bytes=nbytes)
            check(lib.splade_trunk_train_fwd(ctypes.byref(args), ctypes.byref(saves),
                                             torch.cuda.current_stream(dev).cuda_stream), "splade_trunk_train_fwd")
            row_of, valid = packed_row_index(m32, rows)
            if dense:
                re = rows
                hidden = sv["out"][:rows].view(B, S, H).clone()
            else:
                # ragged batch: the backward works on the first `re` rows only (real tokens rounded up to a GEMM tile);
                # rows [T, re) must be zero in every buffer it reads (the forward left them unwritten or computed them
                # from unwritten rows) - one read-back of the row count instead of zero-filling 6 GB up front
                T = int(sv["plan"][1])
                re = min(rows, max(128, (T + 127) // 128 * 128))
                for k in ("x_attn_in", "x_mlp_in", "xb", "qkv", "ab", "xb2", "gg"):
                    sv[k][:, T:re].zero_()
                sv["x_final"][T:re].zero_()
                sv["out"][T:].zero_()
                hidden = sv["out"][row_of]                  # [B,S,H] bf16, zeros at padding
        ctx.owner, ctx.sv, ctx.dims, ctx.dense, ctx.re = owner, sv, (B, S, H, I, nh, L, rows), dense, re
        ctx.ids, ctx.row_of, ctx.valid = ids, row_of, valid
        ctx.n_params = len(params)
        return hidden

    @staticmethod
    def backward(ctx, d_hidden):
        owner, sv = ctx.owner, ctx.sv
        B, S, H, I, nh, L, rows_cap = ctx.dims
        rows = ctx.re            # rows the backward touches (all of them for a dense batch)
        cfg = owner.cfg
        dev = d_hidden.device
        c = owner.weights(dev)
        eps = float(cfg.norm_eps)
        row_of, valid = ctx.row_of, ctx.valid
        flat_rows = row_of.reshape(-1)
        scratch = rows_cap       # packed_row_index sends padding positions to row `rows_cap`
        window = int(cfg.local_attention) // 2

        dense = ctx.dense

        def pack(padded: torch.Tensor) -> torch.Tensor:
            """[B,S,C] -> [rows,C]: padding positions land on a scratch row that is cut off."""
            if dense:
                return padded.reshape(rows, -1)
            out = torch.zeros((scratch + 1, padded.shape[-1]), dtype=padded.dtype, device=dev)
            out.index_copy_(0, flat_rows, padded.reshape(B * S, -1))
            return out[:rows]

        def unpack(packed: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
            """[rows,C] -> [B,S,C] with zeros at padding positions (``splade_trunk_gather_rows``; a view when dense)."""
            if dense:
                return packed.view(B, S, -1)
            C = packed.shape[1]
            if out is None:
                out = torch.empty((B, S, C), dtype=packed.dtype, device=dev)
            check(load_library().splade_trunk_gather_rows(packed.data_ptr(), flat_rows.data_ptr(), B * S,
                                                          C * packed.element_size(), rows, out.data_ptr(),
                                                          torch.cuda.current_stream(dev).cuda_stream),
                  "splade_trunk_gather_rows")
            return out

        with torch.cuda.device(dev), torch.no_grad():
            # The loop below is host-bound (22 layers x ~25 launches): outputs are allocated once and reused, and the
            # library is called directly with cached workspaces instead of through the allocating wrappers.
            lib = load_library()
            stream = torch.cuda.current_stream(dev).cuda_stream
            bf = lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)  # noqa: E731
            dx16, dgg, h, dh = bf(rows, H), bf(rows, I), bf(rows, 2 * I), bf(rows, 2 * I)
            dxb2, dab, dxb = bf(rows, H), bf(rows, H), bf(rows, H)
            dqkv = (torch.empty if dense else torch.zeros)((rows, 3 * H), dtype=torch.bfloat16, device=dev)
            qkv_pad = dab_pad = None
            if not dense:  # padded staging of q / k / v and dO for the attention backward, reused by every layer
                qkv_pad = torch.empty((B, S, 3 * H), dtype=torch.bfloat16, device=dev)
                dab_pad = torch.empty((B, S, H), dtype=torch.bfloat16, device=dev)
            gemm_ws = torch.empty(256, dtype=torch.uint8, device=dev)
            ln_bytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_LN_BWD, 0, 0, H, 0, 0)
            ln_ws = torch.empty(ln_bytes, dtype=torch.uint8, device=dev)
            dgammas = torch.empty((2 * L + 2, H), dtype=torch.float32, device=dev)

            def gemm(A, W, out):  # out [rows,N] bf16 = A [rows,K] @ W[N,K]^T on the tcgen05 GEMM
                N, K = W.shape
                check(lib.splade_trunk_gemm(_lib.EPI_BF16, A.data_ptr(), rows, W.data_ptr(), N, K, out.data_ptr(), N,
                                            gemm_ws.data_ptr(), 256, stream), "splade_trunk_gemm")

            def ln_bwd(dy, x, gamma, slot, dx_out, accumulate, dx16_out=None):
                dgam = dgammas[slot]
                check(lib.splade_trunk_ln_bwd(dy.data_ptr(), int(dy.dtype == torch.float32), x.data_ptr(),
                                              gamma.data_ptr(), eps, rows, H, dx_out.data_ptr(), int(accumulate),
                                              None if dx16_out is None else dx16_out.data_ptr(),
                                              dgam.data_ptr(), ln_ws.data_ptr(), ln_bytes, stream),
                      "splade_trunk_ln_bwd")
                return dgam

            # ---- final_norm
            dy = pack(d_hidden.to(torch.bfloat16)).contiguous()
            dx = torch.empty((rows, H), dtype=torch.float32, device=dev)
            g_final = ln_bwd(dy, sv["x_final"][:rows], c["final_norm"], 2 * L, dx, False, dx16)
            # visibility of keys for the attention backward (padded layout); padding QUERY rows see everything, so
            # that their softmax stays finite (their output gradient is zero: they contribute nothing)
            all_valid = dense or bool(valid.all())
            mask_global = mask_local = None
            if not all_valid or any(not bool(c["layer_array"][l].is_global) for l in range(L)):
                key_ok = valid[:, None, None, :]
                q_pad = ~valid[:, None, :, None]
                pos = torch.arange(S, device=dev)
                near = (pos[:, None] - pos[None, :]).abs() <= window
                mask_global = None if all_valid else (key_ok | q_pad)
                mask_local = (key_ok & near[None, None]) | q_pad
            layer_grads = [None] * L
            for l in range(L - 1, -1, -1):
                w = c["layers"][l]
                t = owner.transposed(dev, l)
                g = {}
                xb2_l, gg_l, ab_l, xb_l = sv["xb2"][l][:rows], sv["gg"][l][:rows], sv["ab"][l][:rows], sv["xb"][l][:rows]
                # ---- MLP block: x_out = x_mid + Wo2(gelu(a) * g), [a | g] = Wi(mlp_norm(x_mid))
                # (dx16 = bf16(dx) comes out of the LayerNorm backward that produced dx)
                gemm(dx16, t["wo2_t"], dgg)                                           # [rows,I]
                g["wo2"] = _mm_f32(dx16.t(), gg_l)                                    # [H,I]
                gemm(xb2_l, w["wi"], h)                                               # [rows,2I] recomputed
                check(lib.splade_trunk_geglu_bwd(h.data_ptr(), dgg.data_ptr(), rows, I, dh.data_ptr(), stream),
                      "splade_trunk_geglu_bwd")
                gemm(dh, t["wi_t"], dxb2)                                             # [rows,H]
                g["wi"] = _mm_f32(dh.t(), xb2_l)                                      # [2I,H]
                g["mlp_norm"] = ln_bwd(dxb2, sv["x_mlp_in"][l][:rows], w["mlp_norm"], 2 * l + 1, dx, True, dx16)
                # ---- attention block: x_mid = x_in + Wo(attention(rope(Wqkv(attn_norm(x_in)))))
                gemm(dx16, t["wo_t"], dab)                                            # [rows,H]
                g["wo"] = _mm_f32(dx16.t(), ab_l)
                is_global = bool(c["layer_array"][l].is_global)
                qkv = unpack(sv["qkv"][l][:rows], qkv_pad).view(B, S, 3, nh, 64)
                q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))            # [B,nh,S,64]
                d_o = unpack(dab, dab_pad).view(B, S, nh, 64).transpose(1, 2)
                with torch.enable_grad():
                    q_, k_, v_ = (x.detach().requires_grad_(True) for x in (q, k, v))
                    o = F.scaled_dot_product_attention(q_, k_, v_, attn_mask=mask_global if is_global else mask_local,
                                                       scale=0.125)
                    dq, dk, dv = torch.autograd.grad(o, (q_, k_, v_), d_o)
                cos, sin = owner.rope_tables(dev, S, c["global_theta"] if is_global else c["local_theta"])
                # transpose of the RoPE rotation + interleave into the rows of the Wqkv output gradient: one kernel
                _dqkv_pack(dq, dk, dv, None if dense else row_of, cos, sin, rows, H, out=dqkv)
                del dq, dk, dv, o
                gemm(dqkv, t["wqkv_t"], dxb)                                          # [rows,H]
                g["wqkv"] = _mm_f32(dqkv.t(), xb_l)                                   # [3H,H]
                if w["attn_norm"] is not None:
                    g["attn_norm"] = ln_bwd(dxb, sv["x_attn_in"][l][:rows], w["attn_norm"], 2 * l, dx, True, dx16)
                else:  # layer 0: attn_norm is Identity
                    dx += dxb
                layer_grads[l] = g
            # ---- embeddings: x0 = LayerNorm(tok_embeddings[input_ids]); the tied matrix gets a scatter-add
            tok_of_row = torch.zeros(scratch + 1, dtype=torch.int64, device=dev)
            tok_of_row.index_copy_(0, flat_rows, torch.arange(B * S, device=dev))
            ids_packed = ctx.ids.reshape(-1)[tok_of_row[:rows]]
            e = c["tok_emb"][ids_packed]                                              # [rows,H] fp32
            de = torch.empty_like(e)
            g_emb_norm = ln_bwd(dx, e.contiguous(), c["emb_norm"], 2 * L + 1, de, False)
            d_tok = torch.zeros_like(c["tok_emb"])
            d_tok.index_add_(0, ids_packed, de)   # rows behind the last token / replica rows carry zero gradients
        # gradients in the order of NativeTrunkTrain.parameters_in_order()
        out = [d_tok, g_emb_norm]
        for l in range(L):
            g = layer_grads[l]
            if c["layers"][l]["attn_norm"] is not None:
                out.append(g["attn_norm"])
            out += [g["wqkv"], g["wo"], g["mlp_norm"], g["wi"], g["wo2"]]
        out.append(g_final)
        assert len(out) == ctx.n_params
        return (None, None, None, *out)


class NativeTrunkTrain:
    """``hidden = NativeTrunkTrain(mlm)(input_ids, attention_mask)`` = ``mlm.model(...)[0]`` under autocast(bf16), as a
    bf16 tensor [B,S,H] (zeros at padding positions) connected to the trunk's parameters by autograd."""

    def __init__(self, mlm: torch.nn.Module):
        self.mlm = mlm
        self.cfg = mlm.config
        self.base = NativeTrunk(mlm)
        cfg = self.cfg
        if cfg.hidden_size % 256 != 0 or cfg.hidden_size > 1024:
            raise RuntimeError("splade_b200.NativeTrunkTrain: hidden size must be 256, 512, 768 or 1024")
        for name in ("embedding_dropout", "attention_dropout", "mlp_dropout"):
            if float(getattr(cfg, name, 0.0) or 0.0) != 0.0:
                raise RuntimeError(f"splade_b200.NativeTrunkTrain: config.{name} != 0 is not supported")
        self._t = {}
        self._t_versions = None
        self._rope = {}

    # ------------------------------------------------------------------ parameters / cached copies
    def parameters_in_order(self) -> List[torch.nn.Parameter]:
        m = self.mlm
        ps = [m.model.embeddings.tok_embeddings.weight, m.model.embeddings.norm.weight]
        for layer in m.model.layers:
            if isinstance(layer.attn_norm, torch.nn.LayerNorm):
                ps.append(layer.attn_norm.weight)
            ps += [layer.attn.Wqkv.weight, layer.attn.Wo.weight, layer.mlp_norm.weight, layer.mlp.Wi.weight,
                   layer.mlp.Wo.weight]
        ps.append(m.model.final_norm.weight)
        return ps

    def weights(self, device):
        return self.base._refresh(device)

    def transposed(self, device, layer: int):
        """bf16 transposes of the layer's Linear weights (the K-major operands of the dgrad GEMMs), refreshed with
        the bf16 copies themselves."""
        versions = self.base._versions
        if versions != self._t_versions:
            self._t, self._t_versions = {}, versions
        hit = self._t.get(layer)
        if hit is None:
            w = self.weights(device)["layers"][layer]
            hit = {k + "_t": w[k].t().contiguous() for k in ("wqkv", "wo", "wi", "wo2")}
            self._t[layer] = hit
        return hit

    def rope_tables(self, device, S: int, theta: float):
        key = (str(device), S, theta)
        hit = self._rope.get(key)
        if hit is None:
            inv_freq = 1.0 / (theta ** (torch.arange(0, 64, 2, dtype=torch.float32, device=device) / 64.0))
            ang = torch.arange(S, dtype=torch.float32, device=device)[:, None] * inv_freq[None, :]
            hit = (ang.cos().contiguous(), ang.sin().contiguous())   # [S,32] fp32 each
            self._rope[key] = hit
        return hit

    # ------------------------------------------------------------------ forward
    def __call__(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        if not input_ids.is_cuda:
            raise RuntimeError("splade_b200.NativeTrunkTrain: inputs must be CUDA tensors (no CPU fallback)")
        return _TrunkTrainFunction.apply(self, input_ids, attention_mask, *self.parameters_in_order())


def packed_row_index(mask: torch.Tensor, rows: int):
    """Packed row of every token of a right-padded batch: documents lie back to back, each starting at a multiple of 8
    rows (empty documents take none); padding positions point at row ``rows`` (a zero / scratch row).
    mask [B,S] -> (row_of [B,S] int64, valid [B,S] bool); no device synchronisation."""
    B, S = mask.shape
    lens = (mask != 0).sum(dim=1, dtype=torch.int64)
    l8 = (lens + 7) // 8 * 8
    off = torch.cumsum(l8, 0) - l8
    pos = torch.arange(S, device=mask.device)
    valid = pos[None, :] < lens[:, None]
    row_of = torch.where(valid, off[:, None] + pos[None, :], torch.full_like(off[:, None], rows))
    return row_of, valid


def suffix_mask(attention_mask: torch.Tensor) -> bool:
    """True when every row of the mask is ones followed by zeros (right padding) - one device sync."""
    m = attention_mask != 0
    return bool((m[:, 1:] <= m[:, :-1]).all())


__all__ = ["NativeTrunkTrain", "packed_row_index", "suffix_mask"]
