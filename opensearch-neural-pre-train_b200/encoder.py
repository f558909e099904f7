"""Drop-in ``SPLADEModernBERT`` (reference: src/model/splade_modern.py:19-114).

Same class name, constructor, attributes (``model``, ``config``, ``model_name``, ``relu``), properties, method
signatures and ``state_dict`` keys as the reference, so checkpoints (src/train/cli/train_v33_ddp.py:192-273),
``benchmark/encoders.py:268-273`` and the DDP trainer keep working unchanged.  Only the computation after the
ModernBERT trunk differs: instead of materialising ``logits [B,S,V]`` and running the ATen chain
``relu -> log1p -> * mask -> max(dim=1) / max(dim=-1)``, the trunk's head output ``[B,S,H]`` goes to ONE fused
tcgen05 kernel (``splade_head_fwd``); the backward is the gather-driven sparse kernel (``splade_head_bwd``).

Numerics: the head always computes like the reference under ``torch.autocast("cuda", bfloat16)`` — bf16
operands, fp32 accumulate, logits rounded once to bf16 before relu/log1p/max — whether or not autocast is
active (outside autocast the reference is fp32; the difference is bounded by bf16 rounding).

Inference can additionally run the NATIVE trunk (``trunk.NativeTrunk``): the whole
encoder on packed real tokens in one C-ABI call, 3-12x faster than the transformers trunk on this GPU.  It is
opt-in for ``encode`` (``model.use_native_trunk(True)`` or ``SPLADE_B200_NATIVE_TRUNK=1``) because its bf16
arithmetic differs from the transformers trunk by rounding noise (cosine > 0.99999, top-64 overlap ~99.4 %);
``forward`` (the training path, which needs autograd through the trunk) uses the reference trunk unless
``use_native_trunk(train=True)`` selects ``trunk_train.NativeTrunkTrain`` (native forward, hand-driven backward).
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from transformers import AutoModelForMaskedLM

from .ops import autocast_is_bf16, head_prologue_enabled, prologue_of, splade_head


class SPLADEModernBERT(nn.Module):
    """SPLADE-max model with ModernBERT backbone; B200-native head (see module docstring)."""

    def __init__(self, model_name: str = "skt/A.X-Encoder-base", dropout: float = 0.1):
        super().__init__()
        self.model_name = model_name
        self.model = AutoModelForMaskedLM.from_pretrained(model_name)
        self.config = self.model.config
        self.relu = nn.ReLU()  # kept for attribute / state_dict parity; the fused kernel applies relu itself
        # plain Python state (not parameters / buffers: state_dict keys must not change)
        self._native_trunk = None
        self._native_trunk_enabled = os.environ.get("SPLADE_B200_NATIVE_TRUNK", "0") == "1"
        self._native_train = None
        self._native_train_enabled = os.environ.get("SPLADE_B200_NATIVE_TRUNK_TRAIN", "0") == "1"

    def use_native_trunk(self, enabled: bool = True, train: Optional[bool] = None) -> "SPLADEModernBERT":
        """Route ``encode`` (inference, no autograd) through the native packed-token trunk.  ``train=True``
        additionally routes ``forward`` under autocast(bf16) with autograd through ``NativeTrunkTrain`` (native
        forward, hand-driven backward; opt-in: the reference trunk is the default for training)."""
        self._native_trunk_enabled = bool(enabled)
        if train is not None:
            self._native_train_enabled = bool(train)
        return self

    def _native_trainer(self):
        if getattr(self, "_native_train", None) is None:
            from .trunk_train import NativeTrunkTrain
            self._native_train = NativeTrunkTrain(self.model)
        return self._native_train

    def _native(self):
        if getattr(self, "_native_trunk", None) is None:
            from .trunk import NativeTrunk
            self._native_trunk = NativeTrunk(self.model)
        return self._native_trunk

    @property
    def vocab_size(self) -> int:
        return self.config.vocab_size

    @property
    def hidden_size(self) -> int:
        return self.config.hidden_size

    def _head_hidden(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """Trunk exactly as the reference runs it (transformers ModernBertForMaskedLM.forward up to, but not
        including, ``self.decoder``): ModernBertModel -> head.dense -> GELU -> head.norm."""
        mlm = self.model
        outputs = mlm.model(input_ids=input_ids, attention_mask=attention_mask)
        return mlm.head(outputs[0])

    def _forward_impl(self, input_ids, attention_mask, want_token_weights: bool):
        mlm = self.model
        decoder = mlm.decoder
        # N1: under autocast(bf16) the prediction head (dense -> GELU -> LayerNorm) runs fused in front of the SPLADE
        # head (tcgen05 GEMM + one GELU/LayerNorm pass emitting bf16; fused backward); otherwise - fp32 evaluation, a
        # head with biases, another activation - the transformers module runs as in the reference
        prologue = prologue_of(mlm.head) if head_prologue_enabled() and input_ids.is_cuda else None
        if prologue is not None and autocast_is_bf16():
            if getattr(self, "_native_train_enabled", False) and torch.is_grad_enabled():
                from .trunk_train import suffix_mask
                if suffix_mask(attention_mask):
                    trunk_out = self._native_trainer()(input_ids, attention_mask)
                    return splade_head(trunk_out, decoder.weight, decoder.bias, attention_mask,
                                       want_token_weights=want_token_weights, prologue=prologue)
            trunk_out = mlm.model(input_ids=input_ids, attention_mask=attention_mask)[0]
            return splade_head(trunk_out, decoder.weight, decoder.bias, attention_mask,
                               want_token_weights=want_token_weights, prologue=prologue)
        hidden = self._head_hidden(input_ids, attention_mask)
        return splade_head(hidden, decoder.weight, decoder.bias, attention_mask,
                           want_token_weights=want_token_weights)

    def forward(
        self,
        input_ids: torch.Tensor,
        attention_mask: torch.Tensor,
        token_type_ids: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Returns (sparse_repr [B,V] fp32, token_weights [B,S] fp32); ``token_type_ids`` is ignored."""
        sparse_repr, token_weights = self._forward_impl(input_ids, attention_mask, True)
        return sparse_repr, token_weights

    def encode(self, input_ids: torch.Tensor, attention_mask: torch.Tensor) -> torch.Tensor:
        """Inference shortcut: the reference discards token_weights here (splade_modern.py:96), so the
        per-position reduction is not computed at all."""
        if getattr(self, "_native_trunk_enabled", False) and input_ids.is_cuda and not torch.is_grad_enabled():
            return self._native().encode(input_ids, attention_mask)
        sparse_repr, _ = self._forward_impl(input_ids, attention_mask, False)
        return sparse_repr

    def get_top_k_tokens(self, sparse_repr: torch.Tensor, tokenizer, k: int = 50) -> Dict[str, float]:
        """Top-k tokens of ONE vector [V] (splade_modern.py:99-114).  Uses the CUDA top-k (ties by ascending
        id — torch.topk's tie order is unspecified) when the vector is on the GPU."""
        k = min(k, sparse_repr.shape[0])
        if sparse_repr.is_cuda:
            from .ops import topk
            ids, vals, counts = topk(sparse_repr.detach().float().unsqueeze(0), k)
            n = int(counts[0])
            top_ids, top_vals = ids[0, :n].tolist(), vals[0, :n].tolist()
        else:
            raise RuntimeError("splade_b200: get_top_k_tokens needs a CUDA tensor (no CPU fallback)")
        result = {}
        for val, idx in zip(top_vals, top_ids):
            if val > 0:
                result[tokenizer.decode([idx]).strip()] = val
        return result
