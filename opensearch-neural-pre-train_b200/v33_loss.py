"""Drop-in ``SPLADELossV33`` (reference: src/model/losses.py:14-301).

Same constructor, ``forward`` signature, ``loss_dict`` keys (python floats) and ``get_avg_nonzero``; the
tensor math runs in ``libsplade_b200.so`` (FLOPS + nonzero counts in one pass, q.d score GEMM, fused
InfoNCE / KL-KD / MarginMSE scalar stage, one backward kernel) and ALL monitoring scalars come back in ONE
device->host copy instead of the reference's 13 ``.item()`` syncs (losses.py:284-294).

Extension (north_star): ``gather_negatives=True`` all-gathers the positive document
representations over the data-parallel group so every query sees W*B in-batch negatives; FLOPS terms stay
per-rank.  Off by default = the reference's per-GPU local negatives.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from .ops import V33LossFunction, V33PeerLossFunction, autocast_is_bf16

_SCALAR_KEYS = ("loss", "infonce", "flops_q", "flops_d", "flops_neg", "kd", "margin_mse", "nonzero_q",
                "nonzero_d")


class SPLADELossV33(nn.Module):
    """L = InfoNCE + lambda_q FLOPS(q) + lambda_d FLOPS(d) + lambda_neg FLOPS(neg) [+ KD terms]."""

    def __init__(
        self,
        lambda_q: float = 1e-2,
        lambda_d: float = 3e-3,
        temperature: float = 1.0,
        flops_warmup_steps: int = 20000,
        lambda_kd: float = 0.0,
        kd_temperature: float = 1.0,
        lambda_initial_ratio: float = 0.1,
        lambda_margin_mse: float = 0.0,
        lambda_neg: float = 0.0,
        gather_negatives: Optional[bool] = None,
        process_group=None,
        exchange=None,
    ):
        super().__init__()
        self.lambda_q = lambda_q
        self.lambda_d = lambda_d
        self.temperature = temperature
        self.flops_warmup_steps = flops_warmup_steps
        self.lambda_kd = lambda_kd
        self.kd_temperature = kd_temperature
        self.lambda_initial_ratio = lambda_initial_ratio
        self.lambda_margin_mse = lambda_margin_mse
        self.lambda_neg = lambda_neg if lambda_neg > 0 else lambda_d  # losses.py:50

        if gather_negatives is None:
            gather_negatives = os.environ.get("SPLADE_B200_GATHER_NEGATIVES", "0") == "1"
        self.gather_negatives = bool(gather_negatives)
        self.process_group = process_group
        # a PeerExchange (peer_exchange.py): gathered negatives through NVLink peer memory inside the loss
        # kernels instead of NCCL all-gather / reduce-scatter around them
        self.exchange = exchange
        if exchange is not None:
            self.gather_negatives = True

        # Monitoring state (losses.py:53-55)
        self._avg_nonzero_q = 0.0
        self._avg_nonzero_d = 0.0
        self._count = 0

    def _lambda_schedule(self, step: int, target_lambda: float) -> float:
        """lambda(t) = target * (r0 + (1 - r0) * (t/T)^2), t >= T -> target (losses.py:75-90)."""
        r0 = self.lambda_initial_ratio
        if step >= self.flops_warmup_steps:
            return target_lambda
        ratio = step / max(self.flops_warmup_steps, 1)
        return target_lambda * (r0 + (1.0 - r0) * ratio * ratio)

    def forward(
        self,
        anchor_repr: torch.Tensor,
        positive_repr: torch.Tensor,
        negative_repr: torch.Tensor,
        global_step: int = 0,
        teacher_scores: Optional[torch.Tensor] = None,
        teacher_pos_scores: Optional[torch.Tensor] = None,
        teacher_neg_scores: Optional[torch.Tensor] = None,
        **kwargs,
    ) -> Tuple[torch.Tensor, Dict[str, float]]:
        cur_lambda_q = self._lambda_schedule(global_step, self.lambda_q)
        cur_lambda_d = self._lambda_schedule(global_step, self.lambda_d)
        cur_lambda_neg = self._lambda_schedule(global_step, self.lambda_neg)

        negs = negative_repr if negative_repr.dim() == 3 else negative_repr.unsqueeze(1)
        if teacher_neg_scores is not None and teacher_neg_scores.dim() == 1:
            teacher_neg_scores = teacher_neg_scores.unsqueeze(1)

        docs, label_offset = positive_repr, 0
        use_peer = self.exchange is not None and self.exchange.world > 1
        if self.gather_negatives and not use_peer:
            from .distributed import all_gather_documents
            docs, label_offset = all_gather_documents(positive_repr, self.process_group)

        use_kd = self.lambda_kd > 0 and teacher_scores is not None
        if use_kd and (use_peer or docs.shape[0] != anchor_repr.shape[0]):
            raise RuntimeError("SPLADELossV33: KL-KD teacher_scores [B,B] cannot be combined with gathered negatives")
        use_mmse = (self.lambda_margin_mse > 0 and teacher_pos_scores is not None
                    and teacher_neg_scores is not None)

        cfg = dict(
            label_offset=label_offset,
            # autocast makes torch.mm consume and produce bf16; mirror it bit for bit
            round_bf16=autocast_is_bf16(),
            temperature=float(self.temperature), kd_temperature=float(self.kd_temperature),
            lambda_q=float(cur_lambda_q), lambda_d=float(cur_lambda_d), lambda_neg=float(cur_lambda_neg),
            lambda_kd=float(self.lambda_kd) if use_kd else 0.0,
            lambda_mmse=float(self.lambda_margin_mse) if use_mmse else 0.0,
        )
        if use_peer:
            loss, scalars = V33PeerLossFunction.apply(
                anchor_repr, positive_repr, negs,
                teacher_pos_scores if use_mmse else None,
                teacher_neg_scores if use_mmse else None,
                cfg, self.exchange,
            )
        else:
            loss, scalars = V33LossFunction.apply(
                anchor_repr, docs, negs,
                teacher_scores if use_kd else None,
                teacher_pos_scores if use_mmse else None,
                teacher_neg_scores if use_mmse else None,
                cfg,
            )

        host = scalars.tolist()  # the single device->host sync of this step (16 floats)
        vals = host[: len(_SCALAR_KEYS)]
        if use_peer:
            # scalars[15] carries the exchange window's status word (raw int32 bits): a barrier of this rank that
            # gave up - in this forward or in an earlier backward - must stop the run, not poison it silently
            status = struct.unpack("<i", struct.pack("<f", host[15]))[0]
            if status != 0:
                self.exchange.raise_for_status(status)
        s = dict(zip(_SCALAR_KEYS, vals))
        self._avg_nonzero_q = 0.9 * self._avg_nonzero_q + 0.1 * s["nonzero_q"]
        self._avg_nonzero_d = 0.9 * self._avg_nonzero_d + 0.1 * s["nonzero_d"]
        self._count += 1

        loss_dict = {
            "infonce": s["infonce"],
            "flops_q": s["flops_q"],
            "flops_d": s["flops_d"],
            "flops_neg": s["flops_neg"],
            "lambda_q": cur_lambda_q,
            "lambda_d": cur_lambda_d,
            "lambda_neg": cur_lambda_neg,
            "kd": s["kd"],
            "margin_mse": s["margin_mse"],
            "nonzero_q": s["nonzero_q"],
            "nonzero_d": s["nonzero_d"],
        }
        return loss, loss_dict

    def get_avg_nonzero(self) -> Tuple[float, float]:
        """Return EMA of nonzero activations (query, doc)."""
        return self._avg_nonzero_q, self._avg_nonzero_d
