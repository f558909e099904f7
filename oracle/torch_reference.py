"""The reference path stated with stock torch ops.  TEST / BASELINE INFRASTRUCTURE ONLY.

Two uses, both outside the product package:
  * on the GPU box, run under ``torch.autocast("cuda", bfloat16)`` it IS the reference's training numerics
    (cuBLAS bf16 GEMM + ATen relu/log1p/mul/max chain) — the apples-to-apples oracle for
    bit-exact argmax / top-k ids, and the "fair GPU baseline" the fused kernel has to beat;
  * on the host CPU in fp32 it is the reference's CPU inference path (benchmark/encoders.py:309-318), timed by
    ``bench.py`` as ``cpu_baseline`` / ``--impl reference``.

Statements follow src/model/splade_modern.py:69-88 and src/model/losses.py:57-73,136-181 with the HF decoder
``nn.Linear`` (transformers modeling_modernbert.py:562) written as ``F.linear``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def head_chain(hidden: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, attention_mask: torch.Tensor):
    """hidden [B,S,H], weight [V,H], bias [V], attention_mask [B,S] -> (sparse_repr, argmax, token_weights)."""
    logits = F.linear(hidden, weight, bias)                       # K1
    scores = torch.log1p(torch.relu(logits))                      # K2, K3 (log1p is fp32 under autocast)
    scores = scores * attention_mask.unsqueeze(-1).float()        # K4
    sparse_repr, argmax = scores.max(dim=1)                       # K5
    token_weights = scores.max(dim=-1).values                     # K6
    return sparse_repr, argmax, token_weights


def head_chain_autocast(hidden, weight, bias, attention_mask):
    """The same chain under the trainer's autocast context (train_v33_ddp.py:337). CUDA only."""
    with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
        return head_chain(hidden, weight, bias, attention_mask)


def head_prologue_chain(x: torch.Tensor, dense_weight: torch.Tensor, norm_weight: torch.Tensor, eps: float,
                        autocast: bool = True):
    """transformers' ModernBertPredictionHead, ``norm(act(dense(x)))`` (modeling_modernbert.py:493-502; no biases,
    exact GELU), which the reference runs between trunk and decoder (src/model/splade_modern.py:69-73), with stock
    torch ops -> (dense output, GELU output, LayerNorm output).  Under autocast: bf16 Linear, bf16 GELU, fp32 LayerNorm."""
    if autocast:
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            y = F.linear(x, dense_weight)
            g = F.gelu(y)
            h = F.layer_norm(g, (x.shape[-1],), norm_weight, None, eps)
        return y, g, h
    y = F.linear(x, dense_weight)
    g = F.gelu(y)
    return y, g, F.layer_norm(g, (x.shape[-1],), norm_weight, None, eps)


def flops_reg(rep: torch.Tensor) -> torch.Tensor:
    return (rep.mean(dim=0) ** 2).sum()


def v33_loss(anchor, positive, negative, lam_q, lam_d, lam_neg, temperature=1.0):
    """InfoNCE + FLOPS total of SPLADELossV33 (no KD), k hard negatives as [B,V] or [B,k,V]."""
    in_batch = torch.mm(anchor, positive.t()) / temperature
    if negative.dim() == 3:
        hard = (anchor.unsqueeze(1) * negative).sum(dim=-1) / temperature
        neg_flat = negative.reshape(-1, negative.shape[-1])
    else:
        hard = ((anchor * negative).sum(dim=-1) / temperature).unsqueeze(1)
        neg_flat = negative
    labels = torch.arange(anchor.shape[0], device=anchor.device)
    infonce = F.cross_entropy(torch.cat([in_batch, hard], dim=1), labels)
    return infonce + lam_q * flops_reg(anchor) + lam_d * flops_reg(positive) + lam_neg * flops_reg(neg_flat)


class ReferenceSpladeModule(torch.nn.Module):
    """The reference model's statement (src/model/splade_modern.py:31-88) around a given HF MLM model: used by
    the GPU parity tests as the same-device oracle for the drop-in encoder (the reference package itself does
    not exist on the GPU box)."""

    def __init__(self, hf_mlm_model):
        super().__init__()
        self.model = hf_mlm_model
        self.config = hf_mlm_model.config
        self.relu = torch.nn.ReLU()

    def forward(self, input_ids, attention_mask, token_type_ids=None):
        logits = self.model(input_ids=input_ids, attention_mask=attention_mask).logits
        scores = torch.log1p(self.relu(logits)) * attention_mask.unsqueeze(-1).float()
        return scores.max(dim=1).values, scores.max(dim=-1).values
