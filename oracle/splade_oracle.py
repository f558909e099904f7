"""CPU oracle for the SPLADE-max head + SPLADELossV33 hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference's algorithm.  It may be imported only by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` — never by the
product package (``opensearch-neural-pre-train_b200/``), which has no CPU path at all.

Parity pin: the reference's own tests never touch this path ("parity unpinned" by reference
tests), so the oracle is pinned against outputs of the REFERENCE ITSELF, produced in the build container by
``oracle/make_golden.py`` (which imports /root/reference by file path) and committed under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function below against those vectors.

Reference lines restated:
  head forward        src/model/splade_modern.py:69-88  (+ transformers modeling_modernbert.py:562 decoder)
  head backward       autograd of the above (torch.max(dim) -> first maximal index gets the whole gradient)
  FLOPS               src/model/losses.py:57-73
  lambda schedule     src/model/losses.py:75-90
  MarginMSE           src/model/losses.py:92-134
  InfoNCE             src/model/losses.py:136-181
  total / KD / dict   src/model/losses.py:183-297
  top-k emit          benchmark/encoders.py:320-343 ; src/model/splade_modern.py:99-114
  head prologue       transformers modeling_modernbert.py:493-502 (ModernBertPredictionHead), called from
                      src/model/splade_modern.py:69-73 through self.model(...)  [SURVEY 8(f) N1]
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------
# head prologue: hidden = LayerNorm(GELU(dense(x))), no biases, exact-erf GELU
# ------------------------------------------------------------------------------------------------
def head_prologue(x: np.ndarray, dense_w: np.ndarray, norm_w: np.ndarray, eps: float, mode: str = "bf16"):
    """x [T,H], dense_w [H,H], norm_w [H] -> (y, g, hidden).  mode "fp32": the module as written (float64 arithmetic);
    mode "bf16": what torch.autocast(bf16) computes - operands rounded to bf16, the Linear output rounded to bf16, GELU
    evaluated in fp32 and rounded to bf16, LayerNorm in fp32 on those values (its fp32 result is what the decoder's
    own bf16 cast then rounds)."""
    from math import erf, sqrt
    verf = np.vectorize(erf, otypes=[np.float64])
    x64 = np.asarray(x, dtype=np.float64)
    w64 = np.asarray(dense_w, dtype=np.float64)
    if mode == "bf16":
        x64 = bf16_round(x64.astype(np.float32)).astype(np.float64)
        w64 = bf16_round(w64.astype(np.float32)).astype(np.float64)
    y = x64 @ w64.T
    if mode == "bf16":
        y = bf16_round(y.astype(np.float32)).astype(np.float64)
    g = 0.5 * y * (1.0 + verf(y / sqrt(2.0)))
    if mode == "bf16":
        g = bf16_round(g.astype(np.float32)).astype(np.float64)
    mean = g.mean(axis=-1, keepdims=True)
    var = ((g - mean) ** 2).mean(axis=-1, keepdims=True)
    hidden = (g - mean) / np.sqrt(var + eps) * np.asarray(norm_w, dtype=np.float64)
    return y.astype(np.float32), g.astype(np.float32), hidden.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# bf16 helpers (round-to-nearest-even on the fp32 bit pattern, as torch's .to(torch.bfloat16))
# ------------------------------------------------------------------------------------------------
def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 (ties to even), returned as fp32."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    nan = np.isnan(x)
    if nan.any():
        out[nan] = np.nan
    return out.reshape(x.shape)


# ------------------------------------------------------------------------------------------------
# head forward (splade_modern.py:69-88)
# ------------------------------------------------------------------------------------------------
def head_logits(hidden: np.ndarray, W: np.ndarray, bias: np.ndarray, mode: str) -> np.ndarray:
    """logits[b,s,v] = sum_h hidden[b,s,h] W[v,h] + bias[v].

    mode "fp32": the reference without autocast (benchmark/encoders.py:309-318).
    mode "bf16": the reference under torch.autocast(bf16) (train_v33_ddp.py:337): operands and bias rounded
                 to bf16, fp32 accumulate, ONE rounding of (acc + bias) to bf16.
    """
    B, S, H = hidden.shape
    h = hidden.reshape(B * S, H).astype(np.float32)
    w = W.astype(np.float32)
    b = bias.astype(np.float32)
    if mode == "bf16":
        h, w, b = bf16_round(h), bf16_round(w), bf16_round(b)
    logits = h @ w.T + b[None, :]
    if mode == "bf16":
        logits = bf16_round(logits)
    elif mode != "fp32":
        raise ValueError(mode)
    return logits.reshape(B, S, -1)


def pool_logits(logits: np.ndarray, mask: np.ndarray):
    """splade_modern.py:76-86 on given logits.  Returns (sparse_repr [B,V] f32, argmax [B,V] i64,
    token_weights [B,S] f32).  np.argmax returns the FIRST maximal index, like torch.max(dim)."""
    scores = np.log1p(np.maximum(logits.astype(np.float32), np.float32(0))).astype(np.float32)
    scores = scores * mask.astype(np.float32)[:, :, None]
    rep = scores.max(axis=1)
    arg = scores.argmax(axis=1)
    tokw = scores.max(axis=2)
    return rep.astype(np.float32), arg.astype(np.int64), tokw.astype(np.float32)


def head_forward(hidden, W, bias, mask, mode: str = "bf16"):
    """Chunked over documents so the [B,S,V] tensor never exceeds one document batch slice."""
    B = hidden.shape[0]
    reps, args, tokws = [], [], []
    step = max(1, int(2e8 // max(1, hidden.shape[1] * W.shape[0])))
    for b0 in range(0, B, step):
        lg = head_logits(hidden[b0:b0 + step], W, bias, mode)
        r, a, t = pool_logits(lg, mask[b0:b0 + step])
        reps.append(r)
        args.append(a)
        tokws.append(t)
    return np.concatenate(reps), np.concatenate(args), np.concatenate(tokws)


def near_tie_mask(hidden, W, bias, mask, mode: str = "bf16", ulps: int = 1):
    """[B,V] bool: cells where the reference's best and second-best positions are within `ulps` bf16 ulps
    of each other, i.e. where a different accumulation order may legitimately flip the argmax."""
    B = hidden.shape[0]
    out = []
    for b0 in range(B):
        lg = head_logits(hidden[b0:b0 + 1], W, bias, mode)[0]  # [S,V]
        x = np.maximum(lg, 0) * mask[b0].astype(np.float32)[:, None]
        if x.shape[0] < 2:
            out.append(np.zeros(x.shape[1], dtype=bool))
            continue
        part = np.partition(x, -2, axis=0)
        top, second = part[-1], part[-2]
        ulp = np.maximum(np.abs(top), np.float32(2.0 ** -126)) * np.float32(2.0 ** -7)
        out.append((top - second) <= ulps * ulp)
    return np.stack(out)


# ------------------------------------------------------------------------------------------------
# head backward: autograd of splade_modern.py:69-83
# ------------------------------------------------------------------------------------------------
def head_backward(g_rep, hidden, W, bias, mask, mode: str = "bf16"):
    """Returns (d_hidden [B,S,H], dW [V,H], dbias [V]) in float64-accumulated fp32.

    dlogit[b,s,v] = g_rep[b,v] * [s == argmax[b,v]] * mask[b,s] * [x>0] / (1+x), x = logit (bf16-rounded in
    mode "bf16"); the bf16 casts are straight-through, exactly as autograd treats them."""
    B, S, H = hidden.shape
    V = W.shape[0]
    hid = hidden.astype(np.float32)
    w = W.astype(np.float32)
    if mode == "bf16":
        hid, w = bf16_round(hid), bf16_round(w)
    d_hidden = np.zeros((B, S, H), dtype=np.float64)
    dW = np.zeros((V, H), dtype=np.float64)
    dbias = np.zeros((V,), dtype=np.float64)
    cols = np.arange(V)
    for b in range(B):
        lg = head_logits(hidden[b:b + 1], W, bias, mode)[0]  # [S,V]
        _, arg, _ = pool_logits(lg[None], mask[b:b + 1])
        arg = arg[0]
        x = lg[arg, cols].astype(np.float64)
        m = mask[b][arg].astype(np.float64)
        coef = np.where(x > 0, g_rep[b].astype(np.float64) * m / (1.0 + np.maximum(x, 0.0)), 0.0)  # [V]
        dbias += coef
        np.add.at(d_hidden[b], arg, coef[:, None] * w.astype(np.float64))
        dW += coef[:, None] * hid[b][arg].astype(np.float64)
    return d_hidden.astype(np.float32), dW.astype(np.float32), dbias.astype(np.float32)


# ------------------------------------------------------------------------------------------------
# loss (losses.py)
# ------------------------------------------------------------------------------------------------
def lambda_schedule(step: int, target: float, warmup_steps: int, r0: float) -> float:
    """losses.py:75-90"""
    if step >= warmup_steps:
        return target
    ratio = step / max(warmup_steps, 1)
    return target * (r0 + (1.0 - r0) * ratio * ratio)


def flops(rep: np.ndarray) -> float:
    """losses.py:72-73: sum_j (mean_i w_ij)^2"""
    m = rep.astype(np.float64).mean(axis=0)
    return float((m * m).sum())


def _log_softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def _mm(a: np.ndarray, b: np.ndarray, temperature: float, autocast: bool) -> np.ndarray:
    """torch.mm(a, b.t()) / T.  Under autocast both operands are cast to bf16, the product is bf16, and
    dividing a bf16 tensor by a python float stays bf16."""
    if autocast:
        s = bf16_round(bf16_round(a) @ bf16_round(b).T)
        return bf16_round(s / np.float32(temperature)).astype(np.float64)
    return (a.astype(np.float64) @ b.astype(np.float64).T) / temperature


def v33_loss(q, docs, negs, *, global_step=0, lambda_q=1e-2, lambda_d=3e-3, temperature=1.0,
             flops_warmup_steps=20000, lambda_kd=0.0, kd_temperature=1.0, lambda_initial_ratio=0.1,
             lambda_margin_mse=0.0, lambda_neg=0.0, teacher_scores=None, teacher_pos_scores=None,
             teacher_neg_scores=None, autocast=False, label_offset=0, want_grads=True):
    """SPLADELossV33.forward (losses.py:183-297) and its gradients, float64.

    q [Bq,V]; docs [N,V] (N == Bq in the reference; N = W*Bq with gathered negatives, the positive of query i
    being docs[label_offset+i]); negs [Bq,V] or [Bq,k,V].  FLOPS_d uses the LOCAL positives only.
    Returns (loss, loss_dict, (dq, ddocs, dnegs) | None)."""
    q64 = q.astype(np.float64)
    d64 = docs.astype(np.float64)
    n3 = negs if negs.ndim == 3 else negs[:, None, :]
    n64 = n3.astype(np.float64)
    Bq, V = q64.shape
    N = d64.shape[0]
    k = n64.shape[1]
    lam_neg_target = lambda_neg if lambda_neg > 0 else lambda_d
    lq = lambda_schedule(global_step, lambda_q, flops_warmup_steps, lambda_initial_ratio)
    ld = lambda_schedule(global_step, lambda_d, flops_warmup_steps, lambda_initial_ratio)
    ln = lambda_schedule(global_step, lam_neg_target, flops_warmup_steps, lambda_initial_ratio)
    labels = label_offset + np.arange(Bq)
    local = d64[label_offset:label_offset + Bq]

    # InfoNCE (losses.py:136-181)
    s_in = _mm(q.astype(np.float32), docs.astype(np.float32), temperature, autocast)      # [Bq,N]
    raw_hard = np.einsum("iv,ikv->ik", q64, n64)                                              # [Bq,k] fp32 path
    s_hard = raw_hard / temperature
    all_scores = np.concatenate([s_in, s_hard], axis=1)
    logp = _log_softmax(all_scores)
    infonce = float(-logp[np.arange(Bq), labels].mean())

    fq, fd, fn = flops(q64), flops(local), flops(n64.reshape(Bq * k, V))
    loss = infonce + lq * fq + ld * fd + ln * fn

    kd = 0.0
    use_kd = lambda_kd > 0 and teacher_scores is not None
    if use_kd:  # losses.py:239-253
        st = _mm(q.astype(np.float32), docs.astype(np.float32), kd_temperature, autocast)
        t_logp = _log_softmax(teacher_scores.astype(np.float64) / kd_temperature)
        t_p = np.exp(t_logp)
        s_logp = _log_softmax(st)
        kd = float((t_p * (t_logp - s_logp)).sum() / Bq)
        loss += lambda_kd * kd

    mmse = 0.0
    use_mm = lambda_margin_mse > 0 and teacher_pos_scores is not None and teacher_neg_scores is not None
    if use_mm:  # losses.py:92-134
        sp = (q64 * local).sum(-1)                                   # [Bq]
        tneg = teacher_neg_scores.astype(np.float64).reshape(Bq, k)
        diff = (sp[:, None] - raw_hard) - (teacher_pos_scores.astype(np.float64)[:, None] - tneg)
        mmse = float((diff ** 2).mean())
        loss += lambda_margin_mse * mmse

    nz_q = float((q64 > 0).sum(-1).mean())
    nz_d = float((local > 0).sum(-1).mean())
    loss_dict = {"infonce": infonce, "flops_q": fq, "flops_d": fd, "flops_neg": fn, "lambda_q": lq,
                 "lambda_d": ld, "lambda_neg": ln, "kd": kd, "margin_mse": mmse, "nonzero_q": nz_q,
                 "nonzero_d": nz_d}
    if not want_grads:
        return loss, loss_dict, None

    # gradients (bf16 casts are straight-through)
    p = np.exp(logp)
    p[np.arange(Bq), labels] -= 1.0
    g_all = p / Bq / temperature
    g_in, g_hard = g_all[:, :N].copy(), g_all[:, N:].copy()
    if use_kd:
        g_in += lambda_kd * (np.exp(s_logp) - t_p) / Bq / kd_temperature
    if use_mm:
        c = 2.0 * diff / (Bq * k)
        g_hard += lambda_margin_mse * (-c)
        g_in[np.arange(Bq), labels] += lambda_margin_mse * c.sum(-1)
    dq = g_in @ d64 + np.einsum("ik,ikv->iv", g_hard, n64) + 2.0 * lq * q64.mean(0)[None, :] / Bq
    dd = g_in.T @ q64
    dd[label_offset:label_offset + Bq] += 2.0 * ld * local.mean(0)[None, :] / Bq
    dn = g_hard[:, :, None] * q64[:, None, :] + 2.0 * ln * n64.reshape(Bq * k, V).mean(0)[None, None, :] / (Bq * k)
    if negs.ndim == 2:
        dn = dn[:, 0, :]
    return loss, loss_dict, (dq, dd, dn)


# ------------------------------------------------------------------------------------------------
# top-k emit (benchmark/encoders.py:320-343)
# ------------------------------------------------------------------------------------------------
def topk_emit(rep: np.ndarray, k: Optional[int], blocklist: Optional[np.ndarray] = None):
    """Per row: keep weight > 0 and not block-listed, in ascending id order; if more than k survive keep the k
    largest by a STABLE descending sort (python's sorted(..., reverse=True) keeps ties in original =
    ascending-id order).  Returns a list of (ids int64[], weights f32[]) ordered (weight desc, id asc)."""
    out = []
    for row in rep:
        ids = np.nonzero(row > 0)[0]
        if blocklist is not None:
            ids = ids[blocklist[ids] == 0]
        w = row[ids]
        order = np.lexsort((ids, -w.astype(np.float64)))  # primary: weight desc, secondary: id asc
        if k is not None:
            order = order[:k]
        out.append((ids[order].astype(np.int64), w[order].astype(np.float32)))
    return out
