"""The reference's training micro-step (src/train/cli/train_v33_ddp.py:337-374) as a reusable function.

The reference trainer stays the caller in production (it only needs ``SPLADEModernBERT`` / ``SPLADELossV33``
swapped); this module restates its inner loop so that the hot path can be exercised,
tested and benchmarked without the reference's data pipeline: three encoder calls (query, positive,
negatives) and one loss call inside ``torch.autocast("cuda", bfloat16)``, ``(loss / accum).backward()``,
then — every ``accum`` micro-steps — ``clip_grad_norm_(1.0)``, optimizer and scheduler step.

Trainer hygiene the reference lacks is available behind explicit flags and OFF by
default: ``no_sync`` on non-final accumulation micro-steps (the reference all-reduces 597 MB on every one), and
``enable_bf16_gradient_compression`` (the gradient all-reduce, 597 MB fp32 per step of which the tied
decoder/embedding matrix alone is 153.6 MB, travels as bf16: half the NVLink bytes of the one collective the
reference's DDP step has).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Dict, Optional

import torch


@dataclass
class StepConfig:
    gradient_accumulation_steps: int = 1
    gradient_clip: float = 1.0
    skip_allreduce_on_accumulation: bool = False  # N3: DDP.no_sync() on non-final micro-steps
    autocast_dtype: torch.dtype = torch.bfloat16
    # N3, opt-in: ONE encoder call for queries, positives and negatives (right-padded to the longest of the three and
    # concatenated along the batch) instead of three.  Documents are independent in the encoder, so the sparse vectors
    # are the same; what changes is the cost: with the packed native trunks only real tokens are processed, so the
    # padding added here is free, while the GEMMs see one large batch and the host drives one layer loop instead of
    # three.  (Pointless with the transformers trunk, which computes the padding.)
    fuse_encoder_calls: bool = False


def enable_bf16_gradient_compression(ddp_model) -> None:
    """N3, opt-in: DDP casts every gradient bucket to bf16 for the all-reduce and back (PyTorch's stock
    ``bf16_compress_hook``); call once after wrapping the model (src/train/cli/train_v33_ddp.py:539-544 wraps it).
    Parameters and optimizer state stay fp32; only the averaged gradient carries bf16 rounding (relative 2^-9)."""
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    ddp_model.register_comm_hook(state=None, hook=default_hooks.bf16_compress_hook)


def train_micro_step(model, loss_fn, optimizer, scheduler, batch: Dict[str, torch.Tensor], micro_idx: int,
                     global_step: int, cfg: StepConfig, device) -> tuple:
    """One micro-step of train_epoch.  Returns (loss tensor, loss_dict, new global_step)."""
    q_ids = batch["query_input_ids"].to(device, non_blocking=True)
    q_mask = batch["query_attention_mask"].to(device, non_blocking=True)
    p_ids = batch["positive_input_ids"].to(device, non_blocking=True)
    p_mask = batch["positive_attention_mask"].to(device, non_blocking=True)
    n_ids = batch["negative_input_ids"].to(device, non_blocking=True)
    n_mask = batch["negative_attention_mask"].to(device, non_blocking=True)
    num_negatives = batch.get("num_negatives", 1)
    t_pos = batch.get("teacher_pos_scores")
    t_neg = batch.get("teacher_neg_scores")
    t_pos = t_pos.to(device) if t_pos is not None else None
    t_neg = t_neg.to(device) if t_neg is not None else None

    accum = cfg.gradient_accumulation_steps
    last_micro = (micro_idx + 1) % accum == 0
    sync_ctx = contextlib.nullcontext()
    if cfg.skip_allreduce_on_accumulation and not last_micro and hasattr(model, "no_sync"):
        sync_ctx = model.no_sync()

    with sync_ctx:
        with torch.autocast(device_type="cuda", dtype=cfg.autocast_dtype):
            if cfg.fuse_encoder_calls:
                ids, mask = concat_padded((q_ids, p_ids, n_ids), (q_mask, p_mask, n_mask))
                rep, _ = model(ids, mask)
                anchor_repr, positive_repr, negative_repr = rep.split(
                    [q_ids.shape[0], p_ids.shape[0], n_ids.shape[0]], dim=0)
            else:
                anchor_repr, _ = model(q_ids, q_mask)
                positive_repr, _ = model(p_ids, p_mask)
                negative_repr, _ = model(n_ids, n_mask)
            if num_negatives > 1:
                negative_repr = negative_repr.view(anchor_repr.shape[0], num_negatives, -1)
            loss, loss_dict = loss_fn(anchor_repr=anchor_repr, positive_repr=positive_repr,
                                      negative_repr=negative_repr, global_step=global_step,
                                      teacher_pos_scores=t_pos, teacher_neg_scores=t_neg)
        (loss / accum).backward()

    if last_micro:
        exchange = getattr(loss_fn, "exchange", None)
        if exchange is not None and exchange.world > 1:
            exchange.check()  # a backward barrier that gave up must not reach the optimizer
        torch.nn.utils.clip_grad_norm_(model.parameters(), cfg.gradient_clip)
        optimizer.step()
        if scheduler is not None:
            scheduler.step()
        optimizer.zero_grad()
        global_step += 1
    return loss, loss_dict, global_step


def concat_padded(ids_list, mask_list, pad_id: int = 0):
    """Right-pads every (ids, mask) pair to the longest sequence length among them and concatenates along the batch."""
    S = max(t.shape[1] for t in ids_list)
    ids = [torch.nn.functional.pad(t, (0, S - t.shape[1]), value=pad_id) for t in ids_list]
    mask = [torch.nn.functional.pad(t, (0, S - t.shape[1]), value=0) for t in mask_list]
    return torch.cat(ids, dim=0), torch.cat(mask, dim=0)


def synthetic_triplet_batch(batch_size: int, q_len: int, d_len: int, num_negatives: int, vocab_size: int,
                            pad_id: int, seed: int, ragged: bool = True) -> Dict[str, torch.Tensor]:
    """Collator-shaped batch (src/train/data/dataloader.py:46-164) of random token ids: right padded to the
    longest member, int64, ``negative_*`` flattened to [B*k, S] with ``num_negatives``."""
    g = torch.Generator().manual_seed(seed)

    def side(n, length):
        ids = torch.randint(6, max(7, vocab_size - 1), (n, length), generator=g)
        if ragged:
            lens = torch.randint(max(1, length // 8), length + 1, (n,), generator=g)
            lens[0] = length
        else:
            lens = torch.full((n,), length)
        mask = (torch.arange(length)[None, :] < lens[:, None]).long()
        return torch.where(mask.bool(), ids, torch.full_like(ids, pad_id)), mask

    q, qm = side(batch_size, q_len)
    p, pm = side(batch_size, d_len)
    n, nm = side(batch_size * num_negatives, d_len)
    return {"query_input_ids": q, "query_attention_mask": qm, "positive_input_ids": p,
            "positive_attention_mask": pm, "negative_input_ids": n, "negative_attention_mask": nm,
            "num_negatives": num_negatives}
