"""CPU: host-side mirror of the reference interface (no kernels are called)."""
import inspect
import math
import os

import pytest
import torch

from tests.model_util import tiny_modernbert_dir


def test_lambda_schedule_and_ctor_match_reference_semantics():
    import splade_b200
    fn = splade_b200.SPLADELossV33()
    assert (fn.lambda_q, fn.lambda_d, fn.temperature, fn.flops_warmup_steps) == (1e-2, 3e-3, 1.0, 20000)
    assert fn.lambda_neg == fn.lambda_d                      # losses.py:50
    assert splade_b200.SPLADELossV33(lambda_neg=0.02).lambda_neg == 0.02
    assert fn._lambda_schedule(0, 0.01) == pytest.approx(0.001)
    assert fn._lambda_schedule(10000, 0.01) == pytest.approx(0.00325)
    assert fn._lambda_schedule(20000, 0.01) == 0.01 and fn._lambda_schedule(10 ** 9, 0.01) == 0.01
    assert splade_b200.SPLADELossV33(flops_warmup_steps=0)._lambda_schedule(0, 0.5) == 0.5
    assert fn.get_avg_nonzero() == (0.0, 0.0)
    assert len(list(fn.parameters())) == 0 and len(list(fn.buffers())) == 0   # nothing to checkpoint


def test_signatures_mirror_the_reference():
    import splade_b200
    sig = inspect.signature(splade_b200.SPLADEModernBERT.forward)
    assert list(sig.parameters)[1:] == ["input_ids", "attention_mask", "token_type_ids"]
    sig = inspect.signature(splade_b200.SPLADEModernBERT.__init__)
    assert sig.parameters["model_name"].default == "skt/A.X-Encoder-base" and sig.parameters["dropout"].default == 0.1
    sig = inspect.signature(splade_b200.SPLADELossV33.forward)
    assert list(sig.parameters)[1:8] == ["anchor_repr", "positive_repr", "negative_repr", "global_step",
                                         "teacher_scores", "teacher_pos_scores", "teacher_neg_scores"]
    sig = inspect.signature(splade_b200.SPLADELossV33.__init__)
    assert list(sig.parameters)[1:10] == ["lambda_q", "lambda_d", "temperature", "flops_warmup_steps", "lambda_kd",
                                          "kd_temperature", "lambda_initial_ratio", "lambda_margin_mse", "lambda_neg"]


def test_state_dict_keys_and_attributes_are_the_reference_ones():
    """Checkpoints are model.module.state_dict() <-> load_state_dict (train_v33_ddp.py:210-211, 252-255): the
    drop-in adds no parameter or buffer and keeps `self.model` = the HF ModernBertForMaskedLM."""
    import splade_b200
    from transformers import AutoModelForMaskedLM
    d = tiny_modernbert_dir(vocab_size=512, hidden=64, layers=2, heads=2)
    m = splade_b200.SPLADEModernBERT(model_name=d)
    hf = AutoModelForMaskedLM.from_pretrained(d)
    assert set(m.state_dict().keys()) == {"model." + k for k in hf.state_dict().keys()}
    assert m.vocab_size == 512 and m.hidden_size == 64 and m.model_name == d
    assert isinstance(m.relu, torch.nn.ReLU) and m.config is m.model.config
    assert m.model.decoder.weight.data_ptr() == m.model.model.embeddings.tok_embeddings.weight.data_ptr()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4, dtype=torch.long))


def test_shard_range_covers_everything_once():
    import splade_b200
    for n in (0, 1, 7, 64, 1000, 1_000_000):
        for w in (1, 2, 3, 8):
            spans = [splade_b200.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == math.ceil(n / w) if n else True
    assert splade_b200.shard_range(1_000_000, 3, 8) == (375000, 500000)     # BASELINE config 5: 125 000 / GPU


def test_synthetic_batch_has_the_collator_contract():
    from splade_b200.training import synthetic_triplet_batch
    b = synthetic_triplet_batch(4, 16, 32, 3, vocab_size=1000, pad_id=999, seed=0)
    assert b["query_input_ids"].shape == (4, 16) and b["negative_input_ids"].shape == (12, 32)
    assert b["num_negatives"] == 3 and b["query_input_ids"].dtype == torch.int64
    m = b["positive_attention_mask"]
    assert bool((m[:, :-1] >= m[:, 1:]).all())                              # right padded
    assert bool((b["positive_input_ids"][m == 0] == 999).all())


def test_blocklist_from_tokenizer_contract():
    import splade_b200

    class Tok:
        vocab_size = 9
        cls_token_id, sep_token_id, pad_token_id, unk_token_id, bos_token_id, eos_token_id = 5, 3, 9, 2, 0, 1

        def convert_ids_to_tokens(self, ids):
            table = ["<s>", "</s>", "[UNK]", "[SEP]", "[MASK]", "[CLS]", "ab", "##c", "<x"]
            return [table[i] for i in ids]

    block = splade_b200.build_blocklist(Tok(), 10)
    assert block.tolist() == [1, 1, 1, 1, 1, 1, 0, 0, 1, 1]


def test_c_emitter_json_matches_python_dicts():
    """splade_emit_json is host code: it runs in the CPU suite.  Its output must parse to exactly the dicts
    SparseEmitter.to_dicts builds (rank_features form), to {"<id>": w} without a token table (sparse_vector form),
    and to round(w, 4) when asked for the AWS script's 4 decimals."""
    import json

    import splade_b200
    tokens = ["plain", 'quo"te', "back\\slash", "한국어", "tab\there", "nl\nx", "ctl\x01", "", "last"]
    ids = torch.tensor([[3, 1, 0, 8], [4, 5, 6, 2], [0, -1, -1, -1], [-1, -1, -1, -1]], dtype=torch.int32)
    vals = torch.tensor([[1.4142135, 0.7049465, 0.33333334, 1e-7], [0.5, 0.25, 0.125, 3.0],
                         [0.1, 0, 0, 0], [0, 0, 0, 0]], dtype=torch.float32)
    counts = torch.tensor([4, 4, 1, 0], dtype=torch.int32)
    em = splade_b200.SparseEmitter(None, tokens)
    objs = em.to_json(ids, vals, counts)
    dicts = em.to_dicts(ids, vals, counts)
    parsed = [json.loads(o.decode("utf-8")) for o in objs]
    assert [list(d_.keys()) for d_ in parsed] == [list(d_.keys()) for d_ in dicts]
    import numpy as np
    for g_, w_ in zip(parsed, dicts):   # "%.9g" round-trips every fp32 weight exactly
        assert all(np.float32(g_[k_]) == np.float32(w_[k_]) for k_ in g_)
    assert objs[3] == b"{}" and list(json.loads(objs[0]).keys()) == ["한국어", 'quo"te', "plain", "last"]
    # sparse_vector form: decimal term ids as keys, 4 decimals
    em_ids = splade_b200.SparseEmitter(None, None)
    objs4 = em_ids.to_json(ids, vals, counts, round_decimals=4)
    want = em_ids.to_dicts(ids, vals, counts, round_decimals=4)
    got = [json.loads(o) for o in objs4]
    for g, w in zip(got, want):
        assert list(g.keys()) == list(w.keys())
        assert all(abs(g[k_] - w[k_]) <= 1e-4 for k_ in g)
    assert objs4[0].startswith(b'{"3": 1.4142, "1": 0.7049')


def _bulk_fixture():
    tokens = ["plain", 'quo"te', "back\\slash", "한국어", "tab\there", "nl\nx", "ctl\x01", "", "last"]
    ids = torch.tensor([[3, 1, 0, 8], [4, 5, 6, 2], [0, -1, -1, -1], [-1, -1, -1, -1]], dtype=torch.int32)
    vals = torch.tensor([[1.4142135, 0.7049465, 0.33333334, 1e-7], [0.5, 0.25, 0.125, 3.0],
                         [0.1, 0, 0, 0], [0, 0, 0, 0]], dtype=torch.float32)
    counts = torch.tensor([4, 4, 1, 0], dtype=torch.int32)
    return tokens, ids, vals, counts


def _expand_like_opensearch_py(actions):
    """What opensearch-py's helpers.bulk puts on the wire for `index` actions (helpers/actions.py expand_action +
    JSONSerializer: compact separators, ensure_ascii=False): an action line and a source line per document."""
    import json
    lines = []
    for a in actions:
        lines.append(json.dumps({"index": {"_index": a["_index"], "_id": a["_id"]}}, ensure_ascii=False,
                                separators=(",", ":")))
        lines.append(json.dumps(a["_source"], ensure_ascii=False, separators=(",", ":")))
    return lines


def test_c_emitter_bulk_body_is_what_the_indexer_sends():
    """splade_emit_bulk (host code): the `_bulk` body for the reference's rank_features actions
    (benchmark/indexer.py:112-138) - parsed back line by line it equals the actions' expansion; strings that need
    escaping survive in ids, index name, content and tokens; weights round-trip fp32 exactly."""
    import json

    import numpy as np

    import splade_b200
    tokens, ids, vals, counts = _bulk_fixture()
    em = splade_b200.SparseEmitter(None, tokens)
    doc_ids = ["a", 'b"2', "c\\3", "d 4"]
    contents = ['안녕 "x"\n', "", "tab\there", "plain"]
    dicts = em.to_dicts(ids, vals, counts)
    actions = [{"_index": "bench-한", "_id": d, "_source": {"doc_id": d, "content": c, "sparse_embedding": v}}
               for d, c, v in zip(doc_ids, contents, dicts)]
    body = em.to_bulk(ids, vals, counts, "bench-한", doc_ids,
                      [{"doc_id": d, "content": c} for d, c in zip(doc_ids, contents)])
    assert body.endswith(b"\n")
    got = body.decode("utf-8").split("\n")[:-1]
    want = _expand_like_opensearch_py(actions)
    assert len(got) == len(want) == 8
    for g, w in zip(got, want):
        g, w = json.loads(g), json.loads(w)
        if "sparse_embedding" in w:
            assert list(g) == list(w) and g["doc_id"] == w["doc_id"] and g["content"] == w["content"]
            assert list(g["sparse_embedding"]) == list(w["sparse_embedding"])
            assert all(np.float32(g["sparse_embedding"][t]) == np.float32(w["sparse_embedding"][t])
                       for t in w["sparse_embedding"])
        else:
            assert g == w
    # action lines are byte-identical to opensearch-py's serialisation (no floats in them)
    assert got[0::2] == want[0::2]
    assert got[7] == '{"doc_id":"d 4","content":"plain","sparse_embedding":{}}'


def test_c_emitter_bulk_sparse_vector_format_skips_empty_documents():
    """The AWS script's variant (scripts/neural_sparse_search_aws.py:745-770, 993-1013): "<term id>" keys, weights
    rounded to 4 decimals, terms that round to zero dropped, documents left without a term skipped."""
    import json

    import splade_b200
    _, ids, vals, counts = _bulk_fixture()
    vals = vals.clone()
    vals[2, 0] = 2e-5  # document 2's only term rounds to 0.0000 -> the whole document is skipped
    em = splade_b200.SparseEmitter(None, None)
    sources = [{"title": f"t{i}", "content": f"c{i}", "category": "x"} for i in range(4)]
    body = em.to_bulk(ids, vals, counts, "aws", [str(i) for i in range(4)], sources, round_decimals=4,
                      skip_empty=True)
    lines = [json.loads(x) for x in body.decode().split("\n") if x]
    assert [x["index"]["_id"] for x in lines[0::2]] == ["0", "1"]
    assert lines[1] == {"title": "t0", "content": "c0", "category": "x",
                        "sparse_embedding": {"3": 1.4142, "1": 0.7049, "0": 0.3333}}
    assert lines[3]["sparse_embedding"] == {"4": 0.5, "5": 0.25, "6": 0.125, "2": 3.0}
    want = em.to_dicts(ids, vals, counts, round_decimals=4)
    assert want[2] == {} and want[0] == lines[1]["sparse_embedding"]
    # without skip_empty every document is present; without sources the field stands alone
    body_all = em.to_bulk(ids, vals, counts, "aws", ["0", "1", "2", "3"], None, round_decimals=4)
    tail = body_all.decode().split("\n")
    assert len(tail) == 9 and tail[5] == '{"sparse_embedding":{}}' and tail[7] == '{"sparse_embedding":{}}'
    with pytest.raises(ValueError):
        em.to_bulk(ids, vals, counts, "aws", ["0"], None)


def test_c_emitter_search_bodies():
    """splade_emit_search (host code, query leg): rank_feature bool queries as benchmark/searchers.py:158-185
    builds them, at most max_clauses clauses, nothing for a query without terms, `_msearch` framing on request."""
    import json

    import numpy as np

    import splade_b200
    tokens, ids, vals, counts = _bulk_fixture()
    em = splade_b200.SparseEmitter(None, tokens)
    bodies = em.to_search_bodies(ids, vals, counts, size=7, max_clauses=3)
    assert len(bodies) == 4 and bodies[3] == b""
    q0 = json.loads(bodies[0])
    assert q0["size"] == 7 and list(q0) == ["size", "query"]
    clauses = q0["query"]["bool"]["should"]
    assert [c["rank_feature"]["field"] for c in clauses] == ["sparse_embedding.한국어", 'sparse_embedding.quo"te',
                                                             "sparse_embedding.plain"]
    assert [np.float32(c["rank_feature"]["boost"]) for c in clauses] == [np.float32(v) for v in vals[0, :3].tolist()]
    q1 = json.loads(bodies[1])
    assert q1["query"]["bool"]["should"][2]["rank_feature"]["field"] == "sparse_embedding.ctl\x01"
    ms = b"".join(em.to_search_bodies(ids, vals, counts, msearch_index="bench-한", field="emb")).decode("utf-8")
    lines = ms.split("\n")
    assert lines[-1] == "" and len(lines) == 7  # 3 queries x (header, body), the empty query is absent
    assert [json.loads(x) for x in lines[0:6:2]] == [{"index": "bench-한"}] * 3
    assert json.loads(lines[1])["query"]["bool"]["should"][3]["rank_feature"] == {
        "field": "emb.last", "boost": pytest.approx(1e-7, rel=1e-6)}
    with pytest.raises(ValueError):
        splade_b200.SparseEmitter(None, None).to_search_bodies(ids, vals, counts)


def test_bf16_weight_cache_is_tied_to_the_parameter_object_not_its_address():
    """evaluating several checkpoints in one process frees model A and builds model B whose tied
    matrix very likely lands on A's allocator address with the same version counter; B must never be served A's
    bf16 copy.  Also: an in-place update (optimizer step) invalidates, an untouched parameter hits."""
    from splade_b200 import ops
    ops.clear_weight_cache()
    stale = 0
    for i in range(8):
        w = torch.nn.Parameter(torch.full((64, 32), float(i + 1)))
        with torch.no_grad():
            w.mul_(1.0)                       # same version history for every "model"
        c = ops.cached_bf16(w)
        stale += int(float(c[0, 0]) != float(i + 1))
        assert ops.cached_bf16(w) is c        # second call of the same step: cache hit
        del w, c
    assert stale == 0
    w = torch.nn.Parameter(torch.ones(8, 8))
    a = ops.cached_bf16(w)
    with torch.no_grad():
        w.add_(1.0)                           # what an optimizer step does
    b = ops.cached_bf16(w)
    assert a is not b and float(b[0, 0]) == 2.0
    w.data = torch.full((8, 8), 5.0)          # re-pointed storage (load / .to()): data_ptr changes
    assert float(ops.cached_bf16(w)[0, 0]) == 5.0
    ops.clear_weight_cache()


def test_only_bf16_autocast_is_mirrored(monkeypatch):
    """SPLADELossV33 / splade_head mirror torch.autocast('cuda', bfloat16); fp16 autocast must be refused, not
    silently computed as bf16.  (CUDA autocast cannot be entered without a device: the two torch
    queries are patched.)"""
    from splade_b200 import ops
    assert ops.autocast_is_bf16() is False
    monkeypatch.setattr(torch, "is_autocast_enabled", lambda *a: True)
    monkeypatch.setattr(torch, "get_autocast_dtype", lambda *a: torch.bfloat16)
    assert ops.autocast_is_bf16() is True
    monkeypatch.setattr(torch, "get_autocast_dtype", lambda *a: torch.float16)
    with pytest.raises(RuntimeError, match="autocast dtype"):
        ops.autocast_is_bf16()


def _fake_head_kernels(monkeypatch, calls):
    """Plain-torch stand-ins for the three C-ABI head entry points so that the HOST logic of SpladeHeadFunction
    (deferral of dW across head calls) can be exercised on CPU.  Test infrastructure only."""
    from splade_b200 import ops

    def fwd(h16, w16, b32, m32, want_tokw, pack_mode=None, plan_out=None):
        logits = (h16.float() @ w16.float().t() + b32).to(torch.bfloat16).float()
        act = torch.log1p(torch.relu(logits)) * m32[:, :, None].float()
        rep, arg = act.max(dim=1)
        return rep, arg.to(torch.int32), (act.max(dim=2).values if want_tokw else None)

    def dhidden(g32, rep, argmax, h16, w16, want_d_hidden=True, active=None):
        calls.append("dhidden")
        coef = torch.where(rep > 0, g32 * torch.exp(-rep), torch.zeros_like(rep))
        dh = None
        if want_d_hidden:
            B, S, H = h16.shape
            dh = torch.zeros(B, S, H)
            for b in range(B):
                dh[b].index_add_(0, argmax[b].long(), coef[b][:, None] * w16.float())
        return dh, coef  # the "bitmap" slot carries the coefficients for the fake sweep

    def dw(segments, H, V, accumulate_into=None):
        calls.append(("dw", len(segments)))
        dW, db = torch.zeros(V, H), torch.zeros(V)
        for g32, rep, argmax, h16, coef in segments:
            B = h16.shape[0]
            rows = h16.float()[torch.arange(B)[:, None], argmax.long()]          # [B,V,H]
            dW += (coef[:, :, None] * rows).sum(0)
            db += coef.sum(0)
        return dW, db

    def last(last_seg, segments, w16, want_d_hidden=True, accumulate_into=None):
        g32, rep, argmax, h16, _ = last_seg
        dh, coef_last = dhidden(g32, rep, argmax, h16, w16, want_d_hidden)
        V, H = w16.shape
        full = [(g, r, a, h, dhidden(g, r, a, h, w16, False)[1]) for g, r, a, h, _ in segments]
        for _ in segments:
            calls.pop()                       # the coefficient recomputations above are not kernel calls
        dW, db = dw(full, H, V)
        return dh, dW, db

    monkeypatch.setattr(ops, "_require_cuda", lambda t, n: None)
    monkeypatch.setattr(ops, "head_backward_last", last)
    monkeypatch.setattr(ops, "head_forward", fwd)
    monkeypatch.setattr(ops, "rep_bitmap", lambda rep: torch.ones(1, dtype=torch.int32))
    monkeypatch.setattr(ops, "head_backward_dhidden", dhidden)
    monkeypatch.setattr(ops, "head_backward_dw", dw)


def _torch_head(hidden, W, bias, mask):
    st = lambda t: t + (t.to(torch.bfloat16).float() - t).detach()   # noqa: E731  bf16 rounding, fp32 gradient
    logits = st(hidden) @ st(W).t() + bias
    logits = logits + (logits.to(torch.bfloat16).float() - logits).detach()      # bf16 rounding, straight-through
    return (torch.log1p(torch.relu(logits)) * mask[:, :, None].float()).max(dim=1).values


def test_one_backward_one_dw_sweep_host_logic(monkeypatch):
    """Three head calls sharing the tied decoder matrix, one backward (train_v33_ddp.py:339-343, 364): exactly ONE
    dW sweep over three segments, returned by the last head node the engine runs; a stale forward whose graph is
    never back-propagated is not waited for; gradients equal plain autograd; retain_graph passes work again."""
    from splade_b200 import ops
    calls = []
    _fake_head_kernels(monkeypatch, calls)
    g = torch.Generator().manual_seed(0)
    V, H = 40, 16
    W = torch.nn.Parameter(torch.randn(V, H, generator=g).to(torch.bfloat16).float() * 0.3)
    bias = torch.nn.Parameter(torch.zeros(V))
    hs = [torch.randn(b, s, H, generator=g).to(torch.bfloat16).float().requires_grad_(True)
          for b, s in ((3, 5), (3, 7), (6, 7))]
    ms = [torch.ones(h.shape[0], h.shape[1], dtype=torch.int64) for h in hs]

    stale, _ = ops.splade_head(hs[0].detach(), W, bias, ms[0], round_bias=False)   # never back-propagated, kept alive
    reps = [ops.splade_head(h, W, bias, m, round_bias=False)[0] for h, m in zip(hs, ms)]
    loss = sum((r * r).sum() for r in reps)
    loss.backward(retain_graph=True)
    assert calls.count("dhidden") == 3 and [c for c in calls if c != "dhidden"] == [("dw", 3)]
    got = [W.grad.clone(), bias.grad.clone()] + [h.grad.clone() for h in hs]

    W2, b2 = W.detach().clone().requires_grad_(True), bias.detach().clone().requires_grad_(True)
    hs2 = [h.detach().clone().requires_grad_(True) for h in hs]
    sum((_torch_head(h, W2, b2, m) ** 2).sum() for h, m in zip(hs2, ms)).backward()
    for a, b in zip(got, [W2.grad, b2.grad] + [h.grad for h in hs2]):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)

    # the same graph again (retain_graph): one more sweep over three segments, gradients accumulate to 2x
    calls.clear()
    loss.backward()
    assert [c for c in calls if c != "dhidden"] == [("dw", 3)]
    torch.testing.assert_close(W.grad, 2 * got[0], rtol=1e-5, atol=1e-6)

    # only two of the three calls take part in a backward: one sweep over two segments
    calls.clear()
    W.grad = None
    reps = [ops.splade_head(h, W, bias, m, round_bias=False)[0] for h, m in zip(hs, ms)]
    (reps[0].sum() + reps[2].sum()).backward()
    assert [c for c in calls if c != "dhidden"] == [("dw", 2)]
    del stale

    # deferral switched off: every head backward sweeps on its own (three dense gradients, summed by autograd)
    calls.clear()
    monkeypatch.setenv("SPLADE_B200_DEFER_DW", "0")
    W.grad = None
    sum((ops.splade_head(h, W, bias, m, round_bias=False)[0] ** 2).sum() for h, m in zip(hs, ms)).backward()
    assert [c for c in calls if c != "dhidden"] == [("dw", 1)] * 3
    torch.testing.assert_close(W.grad, got[0], rtol=1e-5, atol=1e-6)


def test_token_weights_are_computed_only_when_read(monkeypatch):
    """forward() must return token_weights, but every reference caller discards it (train_v33_ddp.py:339-343): the
    returned LazyTokenWeights runs the token_weights variant of the kernel on first use only, reports shape / dtype /
    device without touching the device, and then behaves like the eager tensor."""
    from splade_b200 import ops
    calls = []
    _fake_head_kernels(monkeypatch, calls)
    real_fwd = ops.head_forward
    fwd_calls = []
    monkeypatch.setattr(ops, "head_forward", lambda *a, **k: (fwd_calls.append(bool(a[4])), real_fwd(*a, **k))[1])
    g = torch.Generator().manual_seed(1)
    W = torch.randn(30, 16, generator=g) * 0.3
    h = torch.randn(2, 5, 16, generator=g)
    m = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1]])
    rep, tokw = ops.splade_head(h, W, None, m, round_bias=False)
    assert isinstance(tokw, ops.LazyTokenWeights) and not tokw.is_materialized and fwd_calls == [False]
    assert tuple(tokw.shape) == (2, 5) and tokw.dtype == torch.float32 and tokw.device == h.device
    assert fwd_calls == [False]                       # metadata is free
    s = tokw.sum()                                    # first real use
    assert tokw.is_materialized and fwd_calls == [False, True]
    _, eager = ops.splade_head(h, W, None, m, round_bias=False, lazy_token_weights=False)
    assert type(eager) is torch.Tensor and torch.equal(tokw + 0, eager) and float(s) == float(eager.sum())
    assert torch.equal(tokw[:, 3:], eager[:, 3:]) and float(tokw[0, 3]) == 0.0
    n = len(fwd_calls)
    _ = tokw.max(), tokw.cpu().numpy()
    assert len(fwd_calls) == n                        # materialised once
    assert ops.splade_head(h, W, None, m, want_token_weights=False, round_bias=False)[1] is None


def test_packed_row_index_and_suffix_mask():
    """Host logic of the native training trunk: packed row of every token (documents back to back, 8-row aligned,
    empty documents take no rows) and the right-padding check."""
    import torch

    from splade_b200.trunk_train import packed_row_index, suffix_mask
    lens = [5, 0, 8, 13, 1]
    S = 16
    mask = torch.tensor([[1] * n + [0] * (S - n) for n in lens])
    rows = 256
    row_of, valid = packed_row_index(mask, rows)
    expect, off = [], 0
    for n in lens:
        expect.append([off + s if s < n else rows for s in range(S)])
        off += (n + 7) // 8 * 8
    assert row_of.tolist() == expect
    assert valid.tolist() == [[s < n for s in range(S)] for n in lens]
    assert suffix_mask(mask)
    holes = mask.clone()
    holes[0, 2] = 0
    assert not suffix_mask(holes)
    assert suffix_mask(torch.ones(3, 7, dtype=torch.int64))


def test_concat_padded():
    import torch

    from splade_b200.training import concat_padded
    a, am = torch.tensor([[5, 6]]), torch.tensor([[1, 1]])
    b, bm = torch.tensor([[7, 8, 9, 0]]), torch.tensor([[1, 1, 1, 0]])
    ids, mask = concat_padded((a, b), (am, bm), pad_id=3)
    assert ids.tolist() == [[5, 6, 3, 3], [7, 8, 9, 0]] and mask.tolist() == [[1, 1, 0, 0], [1, 1, 1, 0]]


def test_prologue_of_eligibility():
    """Which prediction heads the fused prologue may stand in for (no biases, exact GELU, hidden 256..1024 in steps
    of 256); anything else keeps the transformers module."""
    import torch
    from transformers import ModernBertConfig
    from transformers.models.modernbert.modeling_modernbert import ModernBertPredictionHead

    from splade_b200.ops import prologue_of

    def head(**kw):
        hidden = kw.pop("hidden", 256)
        cfg = ModernBertConfig(vocab_size=100, hidden_size=hidden, intermediate_size=hidden * 3 // 2,
                               num_hidden_layers=2, num_attention_heads=hidden // 64, pad_token_id=99, bos_token_id=0,
                               eos_token_id=1, cls_token_id=0, sep_token_id=1, max_position_embeddings=512,
                               global_attn_every_n_layers=2, local_attention=128, **kw)
        return ModernBertPredictionHead(cfg)

    ok = prologue_of(head())
    assert ok is not None and ok[0].shape == (256, 256) and ok[1].shape == (256,) and ok[2] > 0
    assert prologue_of(head(hidden=768)) is not None
    assert prologue_of(head(hidden=128)) is None                      # not a multiple of 256
    assert prologue_of(head(classifier_bias=True)) is None
    assert prologue_of(head(norm_bias=True)) is None
    assert prologue_of(head(classifier_activation="relu")) is None
    assert prologue_of(torch.nn.Linear(4, 4)) is None
