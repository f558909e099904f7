"""CPU, build container only (skipped where /root/reference is absent, e.g. on the GPU box): the drop-in classes
against the REAL reference modules — identical public signatures, and the reference's own `train_epoch`
(src/train/cli/train_v33_ddp.py:289-448, imported unmodified with the three shims of SURVEY.md section 8(c))
drives our classes all the way to the kernel boundary, where the CPU-only container must raise the
"no CPU fallback" error instead of silently computing something."""
import importlib.util
import inspect
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref():
    sm = _load("ref_splade_modern_t", f"{REF}/src/model/splade_modern.py")
    ls = _load("ref_losses_t", f"{REF}/src/model/losses.py")
    return sm.SPLADEModernBERT, ls.SPLADELossV33


def _public_sig(fn):
    return [(n, p.default, p.kind) for n, p in inspect.signature(fn).parameters.items()]


def test_signatures_equal_the_reference(ref):
    import splade_b200
    RefModel, RefLoss = ref
    for name in ("__init__", "forward", "encode", "get_top_k_tokens"):
        assert _public_sig(getattr(splade_b200.SPLADEModernBERT, name)) == _public_sig(getattr(RefModel, name)), name
    assert _public_sig(splade_b200.SPLADELossV33.forward) == _public_sig(RefLoss.forward)
    assert _public_sig(splade_b200.SPLADELossV33.get_avg_nonzero) == _public_sig(RefLoss.get_avg_nonzero)
    ours = _public_sig(splade_b200.SPLADELossV33.__init__)
    theirs = _public_sig(RefLoss.__init__)
    assert ours[:len(theirs)] == theirs            # same leading parameters; ours appends two opt-in keywords
    assert [n for n, _, _ in ours[len(theirs):]] == ["gather_negatives", "process_group", "exchange"]
    for prop in ("vocab_size", "hidden_size"):
        assert isinstance(getattr(splade_b200.SPLADEModernBERT, prop), property)
        assert isinstance(getattr(RefModel, prop), property)
    # same schedule, same bookkeeping defaults
    a, b = splade_b200.SPLADELossV33(), RefLoss()
    for step in (0, 1, 9999, 20000, 50000):
        assert a._lambda_schedule(step, 0.01) == b._lambda_schedule(step, 0.01)
    assert vars(b).keys() <= vars(a).keys()


def test_reference_train_epoch_reaches_our_kernel_boundary(ref, tmp_path):
    import splade_b200
    from splade_b200.training import synthetic_triplet_batch
    from tests.model_util import tiny_modernbert_dir

    # --- shims that make the reference trainer importable as shipped (gotcha 3)
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    sys.modules.setdefault("sentence_transformers", st)
    for name in ("src", "src.model", "src.train", "src.train.data", "src.train.cli"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(REF, *name.split("."))]
            sys.modules[name] = pkg
    sys.modules["src.train.data"].load_training_data = lambda *a, **k: []
    coll = types.ModuleType("src.train.data.collator")
    coll.create_tokenizer = lambda *a, **k: None
    sys.modules["src.train.data.collator"] = coll
    # the drop-in: the reference modules keep their names, the classes are ours
    sm = types.ModuleType("src.model.splade_modern")
    sm.SPLADEModernBERT = splade_b200.SPLADEModernBERT
    ls = types.ModuleType("src.model.losses")
    ls.SPLADELossV33 = splade_b200.SPLADELossV33
    sys.modules["src.model.splade_modern"], sys.modules["src.model.losses"] = sm, ls
    sys.path.insert(0, REF)
    try:
        trainer = _load("ref_train_v33_ddp_t", f"{REF}/src/train/cli/train_v33_ddp.py")
    finally:
        sys.path.remove(REF)
    assert trainer.SPLADEModernBERT is splade_b200.SPLADEModernBERT
    assert trainer.SPLADELossV33 is splade_b200.SPLADELossV33

    model = splade_b200.SPLADEModernBERT(model_name=tiny_modernbert_dir(vocab_size=512, hidden=64, layers=2, heads=2))
    loss_fn = splade_b200.SPLADELossV33()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)

    class Loader(list):
        sampler = types.SimpleNamespace(set_epoch=lambda e: None)

    loader = Loader([synthetic_triplet_batch(2, 8, 16, 1, 512, 511, seed=0)])
    cfg = types.SimpleNamespace(training=types.SimpleNamespace(gradient_accumulation_steps=1, gradient_clip=1.0,
                                                               log_every_n_steps=50))
    trainer.is_main_process = lambda: False
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        trainer.train_epoch(model, loader, loss_fn, opt, None, cfg, epoch=0, global_step=0,
                            device=torch.device("cpu"))


def test_bulk_body_equals_the_reference_indexers_actions():
    """The reference's own `bulk_actions_sparse` (benchmark/indexer.py:112-138, loaded unmodified; its sibling
    modules, which need opensearch-py / AWS credentials helpers that are not installed, are stubbed) builds the
    actions; `SparseEmitter.to_bulk` must produce the same documents line for line."""
    import json

    import numpy as np

    import splade_b200
    saved = {k: sys.modules.get(k) for k in ("benchmark", "benchmark.config", "benchmark.data_loader",
                                             "benchmark.encoders", "benchmark.index_manager")}
    try:
        pkg = types.ModuleType("benchmark")
        pkg.__path__ = []
        sys.modules["benchmark"] = pkg
        for mod, names in (("benchmark.config", ["BenchmarkConfig"]), ("benchmark.data_loader", ["BenchmarkData"]),
                           ("benchmark.encoders", ["BgeM3Encoder", "NeuralSparseEncoder"]),
                           ("benchmark.index_manager", ["IndexManager"])):
            m = types.ModuleType(mod)
            for n in names:
                setattr(m, n, type(n, (), {}))
            sys.modules[mod] = m
        ix = _load("ref_benchmark_indexer_t", f"{REF}/benchmark/indexer.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    tokens = ["서울", "날씨", "##는", 'q"uote', "back\\slash", "맑음"]
    ids = torch.tensor([[0, 5, 1], [3, 4, -1], [-1, -1, -1]], dtype=torch.int32)
    vals = torch.tensor([[2.5, 1.25, 0.3333333], [0.1, 0.05, 0.0], [0.0, 0.0, 0.0]], dtype=torch.float32)
    counts = torch.tensor([3, 2, 0], dtype=torch.int32)
    em = splade_b200.SparseEmitter(None, tokens)
    dicts = em.to_dicts(ids, vals, counts)
    doc_ids, contents = ["d0", "d1", "d2"], ["서울 날씨는 맑음", 'he said "hi"', ""]
    docs = [ix.EncodedDocument(doc_id=d, content=c, dense_embedding=[], sparse_embedding=v)
            for d, c, v in zip(doc_ids, contents, dicts)]
    actions = list(ix.bulk_actions_sparse(docs, "bench-sparse"))
    body = em.to_bulk(ids, vals, counts, "bench-sparse", doc_ids,
                      [{"doc_id": d, "content": c} for d, c in zip(doc_ids, contents)])
    lines = [json.loads(x) for x in body.decode("utf-8").split("\n") if x]
    assert len(lines) == 2 * len(actions)
    for a, head, src in zip(actions, lines[0::2], lines[1::2]):
        assert head == {"index": {"_index": a["_index"], "_id": a["_id"]}}
        assert list(src) == list(a["_source"])
        assert src["doc_id"] == a["_source"]["doc_id"] and src["content"] == a["_source"]["content"]
        want = a["_source"]["sparse_embedding"]
        assert list(src["sparse_embedding"]) == list(want)
        assert all(np.float32(src["sparse_embedding"][t]) == np.float32(want[t]) for t in want)


def test_search_bodies_equal_the_reference_searchers():
    """The reference's own NeuralSparseSearcher.search (benchmark/searchers.py:130-185, loaded unmodified with its
    opensearch-py / encoder imports stubbed) builds the request body from a pruned query dict;
    `SparseEmitter.to_search_bodies` must produce the same body from the compact top-k buffers."""
    import json

    import numpy as np

    import splade_b200
    names = ("opensearchpy", "benchmark", "benchmark.config", "benchmark.encoders")
    saved = {k: sys.modules.get(k) for k in names}
    try:
        osp = types.ModuleType("opensearchpy")
        osp.OpenSearch = type("OpenSearch", (), {})
        sys.modules["opensearchpy"] = osp
        pkg = types.ModuleType("benchmark")
        pkg.__path__ = []
        sys.modules["benchmark"] = pkg
        for mod, cls in (("benchmark.config", ["BenchmarkConfig"]),
                         ("benchmark.encoders", ["BgeM3Encoder", "NeuralSparseEncoder"])):
            m = types.ModuleType(mod)
            for n in cls:
                setattr(m, n, type(n, (), {}))
            sys.modules[mod] = m
        sr = _load("ref_benchmark_searchers_t", f"{REF}/benchmark/searchers.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    tokens = ["서울", "날씨", "##는", 'q"uote', "맑음", "tie_a", "tie_b"]
    # rows as splade_topk returns them: weight descending, ties by ascending term id
    ids = torch.tensor([[0, 4, 5, 6, 1], [3, 2, -1, -1, -1], [-1, -1, -1, -1, -1]], dtype=torch.int32)
    vals = torch.tensor([[2.5, 1.25, 0.5, 0.5, 0.3333333], [0.1, 0.05, 0, 0, 0], [0, 0, 0, 0, 0]],
                        dtype=torch.float32)
    counts = torch.tensor([5, 2, 0], dtype=torch.int32)
    em = splade_b200.SparseEmitter(None, tokens)
    dicts = em.to_dicts(ids, vals, counts)  # what encode_for_query returns for these queries

    class Encoder:
        class tokenizer:  # noqa: N801
            @staticmethod
            def get_vocab():
                return {t: i for i, t in enumerate(tokens)}

        def __init__(self):
            self.i = -1

        def encode_for_query(self, text, top_k=100):
            self.i += 1
            return dicts[self.i]

    class Client:
        def __init__(self):
            self.calls = []

        def search(self, index, body):
            self.calls.append((index, body))
            return {"hits": {"hits": [], "total": {"value": 0}}, "took": 0}

    client = Client()
    searcher = sr.NeuralSparseSearcher(client, "bench-sparse", Encoder(), top_k=10)
    for q in ("q0", "q1", "q2"):
        searcher.search(q)
    assert len(client.calls) == 2  # the empty query never reaches OpenSearch
    ours = em.to_search_bodies(ids, vals, counts, size=10)
    assert ours[2] == b""
    for (index, want), got in zip(client.calls, ours[:2]):
        got = json.loads(got)
        assert index == "bench-sparse" and list(got) == list(want) and got["size"] == want["size"]
        gc, wc = got["query"]["bool"]["should"], want["query"]["bool"]["should"]
        assert [c["rank_feature"]["field"] for c in gc] == [c["rank_feature"]["field"] for c in wc]
        assert all(np.float32(a["rank_feature"]["boost"]) == np.float32(b["rank_feature"]["boost"])
                   for a, b in zip(gc, wc))


def test_synthetic_batches_follow_the_real_collator():
    """A17: the reference's TripletCollator (src/train/data/dataloader.py:46-164, loaded unmodified) on real text
    with the checkpoint's tokenizer, single- and multi-negative, with teacher scores - and
    `synthetic_triplet_batch` (what the GPU tests and benches feed the path) has the same tensor keys, dtypes,
    shape relations, right padding with the same pad id, and `num_negatives` convention."""
    from transformers import AutoTokenizer

    from splade_b200.training import synthetic_triplet_batch
    tok_dir = f"{REF}/huggingface/v33"
    if not os.path.isdir(tok_dir):
        pytest.skip("tokenizer files not present")
    tok = AutoTokenizer.from_pretrained(tok_dir)
    dl = _load("ref_dataloader_t", f"{REF}/src/train/data/dataloader.py")
    coll = dl.TripletCollator(tok, query_max_length=64, doc_max_length=256)
    single = [{"query": "서울 날씨", "positive": "서울은 오늘 맑고 따뜻합니다.", "negative": "부산 맛집 추천",
               "teacher_pos_score": 0.9, "teacher_neg_score": 0.1},
              {"query": "파이썬 설치 방법을 알려줘", "positive": "파이썬은 공식 홈페이지에서 내려받아 설치합니다. " * 6,
               "negative": "자바", "teacher_pos_score": 0.8, "teacher_neg_score": 0.2}]
    multi = [{"query": d["query"], "positive": d["positive"], "negatives": [d["negative"], "다른 문서", "또 다른 문서"],
              "teacher_pos_score": 0.5, "teacher_neg_scores": [0.1, 0.2, 0.3]} for d in single]
    for raw, k in ((single, 1), (multi, 3)):
        real = coll([dict(d) for d in raw])
        synth = synthetic_triplet_batch(len(raw), real["query_input_ids"].shape[1],
                                        real["positive_input_ids"].shape[1], k, len(tok), tok.pad_token_id, seed=0)
        tensor_keys = [key for key, v in real.items() if torch.is_tensor(v) and not key.startswith("teacher")]
        assert sorted(tensor_keys) == sorted(key for key, v in synth.items() if torch.is_tensor(v))
        assert real["num_negatives"] == synth["num_negatives"] == k
        for key in tensor_keys:
            assert real[key].dtype == synth[key].dtype == torch.int64, key
            assert real[key].shape[0] == synth[key].shape[0], key          # negatives flattened to [B*k, S]
        assert real["negative_input_ids"].shape[0] == len(raw) * k
        for side in ("query", "positive", "negative"):
            for batch in (real, synth):
                m, ids = batch[f"{side}_attention_mask"], batch[f"{side}_input_ids"]
                assert bool((m[:, :-1] >= m[:, 1:]).all())                  # right padded
                assert bool((ids[m == 0] == tok.pad_token_id).all())       # same pad id convention
                assert int(m.sum(1).max()) == m.shape[1]                   # padded to the longest member
        # teacher scores keep the shapes SPLADELossV33.forward documents ([B] and [B] / [B,k])
        assert real["teacher_pos_scores"].shape == (len(raw),)
        assert real["teacher_neg_scores"].shape == ((len(raw),) if k == 1 else (len(raw), k))


def test_blocklist_with_the_real_tokenizer_equals_the_reference_filter():
    """A9: with the checkpoint's tokenizer, `build_blocklist` marks exactly the term ids the reference's emit loop
    drops (benchmark/encoders.py:277-287 special ids, :332 tokens starting with '[' or '<') plus the model row
    beyond the tokenizer's vocabulary: {0..36, 64, 95, 49999}, 40 ids."""
    from transformers import AutoTokenizer

    import splade_b200
    tok_dir = f"{REF}/huggingface/v33"
    if not os.path.isdir(tok_dir):
        pytest.skip("tokenizer files not present")
    tok = AutoTokenizer.from_pretrained(tok_dir)
    V = 50000
    block = splade_b200.build_blocklist(tok, V)
    # the reference's rule, restated from its loop: id kept iff not special and its token does not start with [ or <
    special = {tok.cls_token_id, tok.sep_token_id, tok.pad_token_id, tok.unk_token_id, tok.bos_token_id,
               tok.eos_token_id} - {None}
    lookup = tok.convert_ids_to_tokens(list(range(tok.vocab_size)))
    dropped = {i for i in range(V)
               if i in special or i >= len(lookup) or not lookup[i] or lookup[i].startswith(("[", "<"))}
    assert set(torch.nonzero(block).flatten().tolist()) == dropped
    assert dropped == set(range(37)) | {64, 95, 49999}
    em = splade_b200.SparseEmitter.from_tokenizer(tok, V)
    assert em.token_lookup[100] == lookup[100] and len(em.token_lookup) == tok.vocab_size == 49999
