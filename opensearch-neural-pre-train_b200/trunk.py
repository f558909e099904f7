"""Native variable-length ModernBERT trunk for inference.

``NativeTrunk`` wraps a loaded ``transformers.ModernBertForMaskedLM`` (the ``model`` attribute of the reference's
``SPLADEModernBERT``, src/model/splade_modern.py:38) and runs ``splade_trunk_encode``: embeddings, every encoder
layer, final norm, MLM head prologue and the fused SPLADE head, on the real tokens only, in ONE C-ABI call.  The
module's parameters stay the single source of truth: bf16 copies of the Linear weights are cached here (plain
Python state, not buffers: ``state_dict`` keys do not change) and refreshed when a parameter's version changes.
This class is inference only (no autograd); the training side is ``trunk_train.NativeTrunkTrain``.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import check, load_library


def _require_shape(cfg) -> None:
    H, I, nh = cfg.hidden_size, int(cfg.intermediate_size), cfg.num_attention_heads
    if H % 64 != 0 or H < 256 or nh * 64 != H or I % 128 != 0:
        raise RuntimeError(
            f"splade_b200.NativeTrunk: unsupported ModernBERT shape (hidden {H}, heads {nh}, intermediate {I}); "
            f"needs 64-wide heads, hidden >= 256, intermediate % 128 == 0")
    for name in ("attention_bias", "mlp_bias", "norm_bias", "classifier_bias"):
        if getattr(cfg, name, False):
            raise RuntimeError(f"splade_b200.NativeTrunk: config.{name}=True is not supported")
    if getattr(cfg, "hidden_activation", "gelu") != "gelu" or getattr(cfg, "classifier_activation", "gelu") != "gelu":
        raise RuntimeError("splade_b200.NativeTrunk: only the exact (erf) GELU is supported")


class NativeTrunk:
    def __init__(self, mlm: torch.nn.Module):
        self.mlm = mlm
        self.cfg = mlm.config
        _require_shape(self.cfg)
        self._cache = {}
        self._versions = None

    # ------------------------------------------------------------------ weights
    def _params(self):
        m = self.mlm
        ps = [m.model.embeddings.tok_embeddings.weight, m.model.embeddings.norm.weight, m.model.final_norm.weight,
              m.head.dense.weight, m.head.norm.weight, m.decoder.weight, m.decoder.bias]
        for layer in m.model.layers:
            if isinstance(layer.attn_norm, torch.nn.LayerNorm):
                ps.append(layer.attn_norm.weight)
            ps += [layer.attn.Wqkv.weight, layer.attn.Wo.weight, layer.mlp_norm.weight, layer.mlp.Wi.weight,
                   layer.mlp.Wo.weight]
        return ps

    def _refresh(self, device):
        ps = self._params()
        versions = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps)
        if versions == self._versions and self._cache:
            return self._cache
        m, cfg = self.mlm, self.cfg
        bf = lambda w: w.detach().to(device=device, dtype=torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda w: w.detach().to(device=device, dtype=torch.float32).contiguous()  # noqa: E731
        c = {"tok_emb": f32(m.model.embeddings.tok_embeddings.weight), "emb_norm": f32(m.model.embeddings.norm.weight),
             "final_norm": f32(m.model.final_norm.weight), "head_dense": bf(m.head.dense.weight),
             "head_norm": f32(m.head.norm.weight), "dec_w": bf(m.decoder.weight),
             # autocast hands F.linear a bf16 bias
             "dec_bias": m.decoder.bias.detach().to(device).to(torch.bfloat16).float().contiguous(), "layers": []}
        layer_types = getattr(cfg, "layer_types", None)
        if layer_types is None:  # transformers < 4.56: global attention every n-th layer (modeling_modernbert.py)
            n = int(cfg.global_attn_every_n_layers)
            layer_types = ["full_attention" if i % n == 0 else "sliding_attention" for i in range(len(m.model.layers))]
        layer_types = list(layer_types)
        arr = (_lib.TrunkLayer * len(m.model.layers))()
        for i, layer in enumerate(m.model.layers):
            t = {"attn_norm": f32(layer.attn_norm.weight) if isinstance(layer.attn_norm, torch.nn.LayerNorm) else None,
                 "wqkv": bf(layer.attn.Wqkv.weight), "wo": bf(layer.attn.Wo.weight),
                 "mlp_norm": f32(layer.mlp_norm.weight), "wi": bf(layer.mlp.Wi.weight), "wo2": bf(layer.mlp.Wo.weight)}
            c["layers"].append(t)
            arr[i] = _lib.TrunkLayer(
                attn_norm=None if t["attn_norm"] is None else t["attn_norm"].data_ptr(), wqkv=t["wqkv"].data_ptr(),
                wo=t["wo"].data_ptr(), mlp_norm=t["mlp_norm"].data_ptr(), wi=t["wi"].data_ptr(),
                wo2=t["wo2"].data_ptr(), is_global=int(layer_types[i] == "full_attention"), reserved=0)
        c["layer_array"] = arr
        rp = getattr(cfg, "rope_parameters", None)
        if rp is not None and "full_attention" in rp:
            c["global_theta"] = float(rp["full_attention"]["rope_theta"])
            c["local_theta"] = float(rp["sliding_attention"]["rope_theta"])
        else:  # older configs carry the two bases as plain attributes (huggingface/v33/config.json:17,25)
            c["global_theta"] = float(cfg.global_rope_theta)
            c["local_theta"] = float(getattr(cfg, "local_rope_theta", None) or cfg.global_rope_theta)
        self._cache, self._versions = c, versions
        return c

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def encode(self, input_ids: torch.Tensor, attention_mask: torch.Tensor, debug: bool = False):
        """input_ids / attention_mask [B,S] on the GPU -> sparse_repr [B,V] fp32 (and, with ``debug``, a dict with
        argmax, the packed head input, the packed column descriptors and the plan)."""
        if not input_ids.is_cuda:
            raise RuntimeError("splade_b200.NativeTrunk: inputs must be CUDA tensors (no CPU fallback)")
        lib = load_library()
        cfg, dev = self.cfg, input_ids.device
        c = self._refresh(dev)
        B, S = input_ids.shape
        H, I, V = cfg.hidden_size, int(cfg.intermediate_size), cfg.vocab_size
        ids = input_ids.to(torch.int64).contiguous()
        m32 = attention_mask.to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            rep = torch.empty((B, V), dtype=torch.float32, device=dev)
            argmax = torch.empty((B, V), dtype=torch.int32, device=dev)
            nbytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_ENCODE, B, S, H, V, I)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            hidden_out = meta_out = plan_out = None
            if debug:
                rows = lib.splade_packed_rows(B, S)
                hidden_out = torch.empty((rows, H), dtype=torch.bfloat16, device=dev)
                meta_out = torch.empty((rows,), dtype=torch.int32, device=dev)
                plan_out = torch.zeros((5,), dtype=torch.int32, device=dev)
            args = _lib.TrunkArgs(
                B=B, S=S, H=H, I=I, V=V, n_heads=cfg.num_attention_heads, n_layers=len(c["layers"]),
                local_window=int(cfg.local_attention) // 2, norm_eps=float(cfg.norm_eps),
                global_theta=c["global_theta"], local_theta=c["local_theta"], reserved_f=0.0,
                input_ids=ids.data_ptr(), mask=m32.data_ptr(), tok_emb=c["tok_emb"].data_ptr(),
                emb_norm=c["emb_norm"].data_ptr(), layers=c["layer_array"], final_norm=c["final_norm"].data_ptr(),
                head_dense=c["head_dense"].data_ptr(), head_norm=c["head_norm"].data_ptr(),
                dec_w=c["dec_w"].data_ptr(), dec_bias=c["dec_bias"].data_ptr(), rep=rep.data_ptr(),
                argmax=argmax.data_ptr(),
                hidden_out=None if hidden_out is None else hidden_out.data_ptr(),
                meta_out=None if meta_out is None else meta_out.data_ptr(),
                plan_out=None if plan_out is None else plan_out.data_ptr(),
                workspace=ws.data_ptr(), ws_bytes=nbytes)
            check(lib.splade_trunk_encode(ctypes.byref(args), torch.cuda.current_stream(dev).cuda_stream),
                  "splade_trunk_encode")
        if debug:
            return rep, {"argmax": argmax, "hidden_packed": hidden_out, "meta": meta_out, "plan": plan_out}
        return rep


def trunk_gemm(epi: int, A: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Unit-test entry: out = epilogue(A [rows,K] bf16 @ W^T) (``splade_trunk_gemm``)."""
    lib = load_library()
    rows, K = A.shape
    N = W.shape[0] // 2 if epi == _lib.EPI_GEGLU else W.shape[0]
    dev = A.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty((rows, N), dtype=torch.bfloat16, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_GEMM, 0, 0, 0, 0, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.splade_trunk_gemm(epi, A.data_ptr(), rows, W.data_ptr(), N, K, out.data_ptr(), out.shape[1],
                                    ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream),
              "splade_trunk_gemm")
    return out


def trunk_attention(qkv_packed: torch.Tensor, mask: torch.Tensor, n_heads: int, window: int,
                    rope_theta: float = 0.0):
    """Unit-test entry: packed qkv [rows,3H] bf16 -> (packed attention output [rows,H] bf16, meta [rows] int32);
    ``rope_theta > 0`` rotates q and k inside the kernel."""
    lib = load_library()
    B, S = mask.shape
    rows, H3 = qkv_packed.shape
    H = H3 // 3
    dev = qkv_packed.device
    assert rows == lib.splade_packed_rows(B, S)
    with torch.cuda.device(dev):
        out = torch.zeros((rows, H), dtype=torch.bfloat16, device=dev)
        meta = torch.zeros((rows,), dtype=torch.int32, device=dev)
        nbytes = lib.splade_workspace_bytes(_lib.OP_TRUNK_ATTN, B, S, 0, 1, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        check(lib.splade_trunk_attn(qkv_packed.data_ptr(), mask.to(torch.int32).contiguous().data_ptr(), B, S, H,
                                    n_heads, window, float(rope_theta), out.data_ptr(), meta.data_ptr(),
                                    ws.data_ptr(), nbytes, torch.cuda.current_stream(dev).cuda_stream),
              "splade_trunk_attn")
    return out, meta
