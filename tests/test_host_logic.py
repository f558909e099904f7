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
