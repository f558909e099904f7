"""Top-k pruning + sparse emit with the reference's production semantics, and sharded bulk encoding.

Reference: ``NeuralSparseEncoderV33._encode_batch`` (benchmark/encoders.py:309-345) keeps ``weight > 0``, drops
the special token ids (``:277-287``) and tokens whose string starts with ``[`` or ``<`` (``:332``), and — when
``top_k`` is given and more survive — keeps the ``top_k`` largest by a stable descending sort, i.e. ties go
to the lower term id.  The reference does that per row on the host (``.cpu()`` + python loops); here the
selection is one CUDA kernel over the whole batch (``splade_topk``) and only ``[B,k]`` ids/weights cross
PCIe.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .distributed import shard_range, world_info
from .ops import topk


def build_blocklist(tokenizer, model_vocab_size: int, device=None) -> torch.Tensor:
    """uint8 [V] with 1 for every term id the reference never emits (benchmark/encoders.py:277-287, :332):
    cls/sep/pad/unk/bos/eos ids, tokens starting with '[' or '<', empty tokens, and ids beyond the tokenizer's
    vocabulary (the model has V=50 000 rows, the tokenizer 49 999 entries)."""
    block = torch.zeros(model_vocab_size, dtype=torch.uint8)
    special = {tokenizer.cls_token_id, tokenizer.sep_token_id, tokenizer.pad_token_id, tokenizer.unk_token_id,
               tokenizer.bos_token_id, tokenizer.eos_token_id}
    for tid in special:
        if tid is not None and 0 <= tid < model_vocab_size:
            block[tid] = 1
    n_tok = min(tokenizer.vocab_size, model_vocab_size)
    tokens = tokenizer.convert_ids_to_tokens(list(range(n_tok)))
    for i, tok in enumerate(tokens):
        if not tok or tok.startswith(("[", "<")):
            block[i] = 1
    block[n_tok:] = 1
    return block if device is None else block.to(device)


class SparseEmitter:
    """GPU top-k + host-side dict / bulk-line emit.

    ``token_lookup`` maps term id -> token string (``tokenizer.convert_ids_to_tokens``); without it the emit
    uses the decimal id as key (the ``sparse_vector`` format of scripts/neural_sparse_search_aws.py:753).
    """

    def __init__(self, blocklist: Optional[torch.Tensor] = None, token_lookup: Optional[Sequence[str]] = None):
        self.blocklist = blocklist
        self.token_lookup = token_lookup

    @classmethod
    def from_tokenizer(cls, tokenizer, model_vocab_size: int, device=None) -> "SparseEmitter":
        lookup = tokenizer.convert_ids_to_tokens(list(range(min(tokenizer.vocab_size, model_vocab_size))))
        return cls(build_blocklist(tokenizer, model_vocab_size, device), lookup)

    def select(self, sparse_repr: torch.Tensor, top_k: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """[B,V] fp32 (CUDA) -> ids [B,k] int32 (-1 padded), weights [B,k] fp32, counts [B] int32."""
        block = self.blocklist
        if block is not None and block.device != sparse_repr.device:
            block = self.blocklist = block.to(sparse_repr.device)
        return topk(sparse_repr, top_k, block)

    def to_dicts(self, ids: torch.Tensor, vals: torch.Tensor, counts: torch.Tensor,
                 round_decimals: Optional[int] = None) -> List[Dict[str, float]]:
        """One device->host copy of the compact buffers, then ``{token: weight}`` per row."""
        ids_h, vals_h, counts_h = ids.cpu().tolist(), vals.cpu().tolist(), counts.cpu().tolist()
        out = []
        for row_ids, row_vals, n in zip(ids_h, vals_h, counts_h):
            d = {}
            for tid, w in zip(row_ids[:n], row_vals[:n]):
                key = self.token_lookup[tid] if self.token_lookup is not None else str(tid)
                if round_decimals is not None:
                    w = round(w, round_decimals)
                    if w <= 0.0:  # scripts/neural_sparse_search_aws.py:753-755
                        continue
                d[key] = w
            out.append(d)
        return out

    def encode_batch(self, sparse_repr: torch.Tensor, top_k: int) -> List[Dict[str, float]]:
        return self.to_dicts(*self.select(sparse_repr, top_k))

    def _token_table(self):
        """Concatenated UTF-8 token bytes + offsets for the C emitter (built once)."""
        if self.token_lookup is None:
            return None
        if getattr(self, "_table", None) is None:
            import numpy as np
            enc = [(t or "").encode("utf-8") for t in self.token_lookup]
            offsets = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(e) for e in enc], out=offsets[1:])
            self._table = (b"".join(enc), offsets)
        return self._table

    def to_json(self, ids: torch.Tensor, vals: torch.Tensor, counts: torch.Tensor,
                round_decimals: Optional[int] = None) -> List[bytes]:
        """One JSON object (bytes) per document — ``{"token": weight, ...}`` or, without a token table,
        ``{"<id>": weight, ...}`` — produced by the C emitter ``splade_emit_json`` from ONE device->host copy of
        the compact buffers (no python object per term)."""
        import ctypes

        import numpy as np

        from ._lib import check, load_library
        lib = load_library()
        ids_h = np.ascontiguousarray(ids.cpu().numpy(), dtype=np.int32)
        vals_h = np.ascontiguousarray(vals.cpu().numpy(), dtype=np.float32)
        counts_h = np.ascontiguousarray(counts.cpu().numpy(), dtype=np.int32)
        B, k = ids_h.shape
        table = self._token_table()
        tok_ptr = off_ptr = None
        n_tok = 0
        if table is not None:
            tok_buf = ctypes.create_string_buffer(table[0], len(table[0]) + 1)
            tok_ptr, off_ptr, n_tok = ctypes.addressof(tok_buf), table[1].ctypes.data, len(table[1]) - 1
        offsets = np.zeros(B + 1, dtype=np.int64)
        need = ctypes.c_size_t(0)
        dec = -1 if round_decimals is None else int(round_decimals)
        cap = B * (2 + k * 40)  # first guess; grown once if tokens are long
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)   # (no zero-fill: the emitter writes every byte it reports)
            rc = lib.splade_emit_json(ids_h.ctypes.data, vals_h.ctypes.data, counts_h.ctypes.data, B, k, tok_ptr,
                                      off_ptr, n_tok, dec, out.ctypes.data, cap, offsets.ctypes.data,
                                      ctypes.byref(need))
            if rc == 0:
                break
            if need.value > cap:
                cap = need.value
                continue
            check(rc, "splade_emit_json")
        else:
            check(rc, "splade_emit_json")
        raw = out[: need.value].tobytes()
        return [raw[offsets[b]:offsets[b + 1]] for b in range(B)]

    def to_bulk(self, ids: torch.Tensor, vals: torch.Tensor, counts: torch.Tensor, index_name: str,
                doc_ids: Sequence[str], sources: Optional[Sequence[Dict[str, object]]] = None,
                field: str = "sparse_embedding", round_decimals: Optional[int] = None,
                skip_empty: bool = False) -> bytes:
        """The body of an OpenSearch ``_bulk`` request for these documents: per document the action line and the
        source line opensearch-py's ``helpers.bulk`` would serialise from the reference's actions
        (benchmark/indexer.py:112-138 ``{"doc_id", "content", "sparse_embedding"}``;
        scripts/neural_sparse_search_aws.py:993-1013 ``{"title", "content", "category", "sparse_embedding"}`` with
        ``skip_empty=True``).  ``sources[i]`` holds the leading members of ``_source`` (serialised here, once per
        document); the per-term object comes from the C emitter ``splade_emit_bulk``."""
        import ctypes
        import json

        import numpy as np

        from ._lib import check, load_library
        lib = load_library()
        ids_h = np.ascontiguousarray(ids.cpu().numpy(), dtype=np.int32)
        vals_h = np.ascontiguousarray(vals.cpu().numpy(), dtype=np.float32)
        counts_h = np.ascontiguousarray(counts.cpu().numpy(), dtype=np.int32)
        B, k = ids_h.shape
        if len(doc_ids) != B or (sources is not None and len(sources) != B):
            raise ValueError(f"to_bulk: {B} documents, {len(doc_ids)} ids, "
                             f"{None if sources is None else len(sources)} sources")

        def pack(chunks):
            enc = [c.encode("utf-8") for c in chunks]
            off = np.zeros(len(enc) + 1, dtype=np.int64)
            np.cumsum([len(e) for e in enc], out=off[1:])
            blob = b"".join(enc)
            return ctypes.create_string_buffer(blob, len(blob) + 1), off

        id_buf, id_off = pack([str(d) for d in doc_ids])
        head_ptr = head_off_ptr = None
        if sources is not None:
            heads = [json.dumps(dict(src), ensure_ascii=False, separators=(",", ":"))[1:-1] for src in sources]
            head_buf, head_off = pack(heads)
            head_ptr, head_off_ptr = ctypes.addressof(head_buf), head_off.ctypes.data
        table = self._token_table()
        tok_ptr = off_ptr = None
        n_tok = 0
        if table is not None:
            tok_buf = ctypes.create_string_buffer(table[0], len(table[0]) + 1)
            tok_ptr, off_ptr, n_tok = ctypes.addressof(tok_buf), table[1].ctypes.data, len(table[1]) - 1
        offsets = np.zeros(B + 1, dtype=np.int64)
        need = ctypes.c_size_t(0)
        dec = -1 if round_decimals is None else int(round_decimals)
        # capacity estimate (so the FIRST call already writes - and may use the emitter's worker threads): exact for
        # the per-document strings, average-based for the terms; an underestimate costs one retry at the exact size
        n_terms = np.minimum(np.maximum(counts_h, 0), k)
        valid = np.arange(k)[None, :] < n_terms[:, None]
        if table is not None:
            tok_len = np.diff(table[1])
            safe = np.clip(ids_h, 0, len(tok_len) - 1)
            key_bytes = int(tok_len[safe][valid].sum())
        else:
            key_bytes = 11 * int(valid.sum())
        per_term = 5 + (dec + 6 if dec >= 0 else 13)
        cap = int(1.05 * (B * (40 + 2 * len(index_name.encode("utf-8")) + len(field.encode("utf-8")))
                          + int(id_off[-1]) + (int(head_off[-1]) + B if sources is not None else 0)
                          + key_bytes + per_term * int(valid.sum()))) + 4096
        rc = 0
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)
            rc = lib.splade_emit_bulk(ids_h.ctypes.data, vals_h.ctypes.data, counts_h.ctypes.data, B, k, tok_ptr,
                                      off_ptr, n_tok, dec, index_name.encode("utf-8"), field.encode("utf-8"),
                                      ctypes.addressof(id_buf), id_off.ctypes.data, head_ptr, head_off_ptr,
                                      int(skip_empty), out.ctypes.data, cap, offsets.ctypes.data, ctypes.byref(need))
            if rc == 0:
                break
            if need.value > cap:
                cap = need.value
                continue
            break
        check(rc, "splade_emit_bulk")
        return out[: need.value].tobytes()

    def to_search_bodies(self, ids: torch.Tensor, vals: torch.Tensor, counts: torch.Tensor, size: int = 10,
                         field: str = "sparse_embedding", max_clauses: int = 64,
                         msearch_index: Optional[str] = None) -> List[bytes]:
        """The request bodies ``NeuralSparseSearcher.search`` sends (benchmark/searchers.py:158-185): one
        ``rank_feature`` should-clause per kept query term, highest weight first, at most ``max_clauses``.  One
        ``bytes`` per query (``b""`` when no term survived - the reference then skips the request); with
        ``msearch_index`` each is the header + body pair of an ``_msearch`` request, so ``b"".join(...)`` is the
        whole multi-search body for the batch (C emitter ``splade_emit_search``)."""
        import ctypes

        import numpy as np

        from ._lib import check, load_library
        lib = load_library()
        table = self._token_table()
        if table is None:
            raise ValueError("to_search_bodies needs a token table (rank_feature fields are named by token string)")
        ids_h = np.ascontiguousarray(ids.cpu().numpy(), dtype=np.int32)
        vals_h = np.ascontiguousarray(vals.cpu().numpy(), dtype=np.float32)
        counts_h = np.ascontiguousarray(counts.cpu().numpy(), dtype=np.int32)
        B, k = ids_h.shape
        tok_buf = ctypes.create_string_buffer(table[0], len(table[0]) + 1)
        offsets = np.zeros(B + 1, dtype=np.int64)
        need = ctypes.c_size_t(0)
        idx = None if msearch_index is None else msearch_index.encode("utf-8")
        # capacity estimate so that the first call already writes (an underestimate costs one retry at the exact size)
        n_terms = np.minimum(np.minimum(np.maximum(counts_h, 0), k), int(max_clauses))
        valid = np.arange(k)[None, :] < n_terms[:, None]
        tok_len = np.diff(table[1])
        key_bytes = int(tok_len[np.clip(ids_h, 0, len(tok_len) - 1)][valid].sum())
        per_clause = 48 + len(field.encode("utf-8")) + 14
        cap = int(1.1 * (B * (64 + (len(idx) if idx else 0)) + key_bytes + per_clause * int(valid.sum()))) + 4096
        rc = 0
        for _ in range(2):
            out = np.empty(max(cap, 1), dtype=np.uint8)
            rc = lib.splade_emit_search(ids_h.ctypes.data, vals_h.ctypes.data, counts_h.ctypes.data, B, k,
                                        ctypes.addressof(tok_buf), table[1].ctypes.data, len(table[1]) - 1,
                                        int(max_clauses), int(size), field.encode("utf-8"), idx,
                                        out.ctypes.data, cap, offsets.ctypes.data, ctypes.byref(need))
            if rc == 0:
                break
            if need.value > cap:
                cap = need.value
                continue
            break
        check(rc, "splade_emit_search")
        raw = out[: need.value].tobytes()
        return [raw[offsets[b]:offsets[b + 1]] for b in range(B)]


@torch.no_grad()
def bulk_encode(model, input_ids: torch.Tensor, attention_mask: torch.Tensor, top_k: int,
                emitter: Optional[SparseEmitter] = None, batch_size: int = 64, group=None,
                autocast_dtype: Optional[torch.dtype] = torch.bfloat16, native_trunk: Optional[bool] = None):
    """Encode this rank's contiguous shard of a corpus (BASELINE config 5) and return compact top-k buffers.

    ``input_ids`` / ``attention_mask`` are the FULL corpus tensors [N,S] (host or device); rank r encodes rows
    ``shard_range(N, r, W)``.  No collective on the data path; callers concatenate the per-rank results in rank
    order.  Returns (ids [n,k] int32, weights [n,k] fp32, counts [n] int32, (lo, hi)).
    ``native_trunk``: True / False pins the native packed-token trunk on / off for this call, None keeps the
    model's own setting (``SPLADEModernBERT.use_native_trunk``).
    """
    if native_trunk is not None and hasattr(model, "use_native_trunk"):
        previous = getattr(model, "_native_trunk_enabled", False)
        model.use_native_trunk(native_trunk)
        try:
            return bulk_encode(model, input_ids, attention_mask, top_k, emitter, batch_size, group, autocast_dtype)
        finally:
            model.use_native_trunk(previous)
    rank, world = world_info(group)
    lo, hi = shard_range(input_ids.shape[0], rank, world)
    emitter = emitter or SparseEmitter()
    device = next(model.parameters()).device
    out_ids, out_vals, out_counts = [], [], []
    for b0 in range(lo, hi, batch_size):
        b1 = min(hi, b0 + batch_size)
        ids_b = input_ids[b0:b1].to(device, non_blocking=True)
        mask_b = attention_mask[b0:b1].to(device, non_blocking=True)
        if autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                rep = model.encode(ids_b, mask_b)
        else:
            rep = model.encode(ids_b, mask_b)
        i, v, c = emitter.select(rep, top_k)
        out_ids.append(i)
        out_vals.append(v)
        out_counts.append(c)
    if not out_ids:
        dev = device
        return (torch.empty((0, top_k), dtype=torch.int32, device=dev),
                torch.empty((0, top_k), dtype=torch.float32, device=dev),
                torch.empty((0,), dtype=torch.int32, device=dev), (lo, hi))
    return torch.cat(out_ids), torch.cat(out_vals), torch.cat(out_counts), (lo, hi)
