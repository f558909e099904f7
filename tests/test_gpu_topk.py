"""-m gpu: batched top-k pruning vs the reference's emit semantics (golden from benchmark/encoders.py) and the
oracle at the query-encode shape (BASELINE config 3)."""
import numpy as np
import pytest
import torch

from oracle import splade_oracle as O

pytestmark = pytest.mark.gpu


def _compare(ids, vals, counts, ref, k):
    ids, vals, counts = ids.cpu().numpy(), vals.cpu().numpy(), counts.cpu().numpy()
    for r, (rid, rw) in enumerate(ref):
        n = int(counts[r])
        assert n == len(rid), (r, n, len(rid))
        assert ids[r, :n].tolist() == rid.tolist(), r          # bit-exact term ids, in order
        assert np.array_equal(vals[r, :n], rw), r              # weights are copied, not recomputed
        assert (ids[r, n:] == -1).all() and (vals[r, n:] == 0).all()


def test_golden_emit(cuda_device, golden):
    import splade_b200
    g = golden("emit")
    rep = torch.from_numpy(g["rep"]).to(cuda_device)
    block = torch.from_numpy(g["blocklist"]).to(cuda_device)
    tokens = g["tokens"].tolist()
    em = splade_b200.SparseEmitter(block, tokens)
    for tag, k in (("5", 5), ("64", 64)):
        dicts = em.encode_batch(rep, k)
        for r, d in enumerate(dicts):
            ref = dict(zip((tokens[i] for i in g[f"ids_{tag}_{r}"]), g[f"w_{tag}_{r}"].tolist()))
            assert d == ref, (tag, r)
    # top_k larger than any row: everything that survives the filters, as the reference's top_k=None
    dicts = em.encode_batch(rep, 512)
    for r, d in enumerate(dicts):
        ref = dict(zip((tokens[i] for i in g[f"ids_all_{r}"]), g[f"w_all_{r}"].tolist()))
        assert d == ref, ("all", r)


@pytest.mark.parametrize("k", [1, 50, 64, 100, 1000])
def test_ties_and_blocklist_vs_oracle(cuda_device, k):
    import splade_b200
    g = torch.Generator().manual_seed(k)
    B, V = 7, 50000
    rep = torch.rand(B, V, generator=g)
    rep = (rep * 64).floor() / 64                       # 64 distinct values -> ~780 ties per value
    rep = rep * (torch.rand(B, V, generator=g) < 0.3)
    rep[1] = 0
    rep[2, :] = 0.25                                    # one value everywhere: pure id order
    rep[3] = rep[3] * (torch.rand(V, generator=g) < 0.0005)  # fewer survivors than k
    block = torch.zeros(V, dtype=torch.uint8)
    block[list(range(37)) + [64, 95, V - 1]] = 1        # the real tokenizer's blocklist
    ids, vals, counts = splade_b200.topk(rep.to(cuda_device), k, block.to(cuda_device))
    _compare(ids, vals, counts, O.topk_emit(rep.numpy(), k, block.numpy()), k)
    ids, vals, counts = splade_b200.topk(rep.to(cuda_device), k, None)
    _compare(ids, vals, counts, O.topk_emit(rep.numpy(), k, None), k)


def test_query_encode_shape_properties(cuda_device):
    """Config 3 (B=512, V=50 000, k=64): sortedness, idempotence, and exact agreement on a row sample."""
    import splade_b200
    g = torch.Generator().manual_seed(0)
    B, V, k = 512, 50000, 64
    x = torch.randn(B, V, generator=g).to(torch.bfloat16).float()
    rep = torch.log1p(torch.relu(x)).to(cuda_device)
    ids, vals, counts = splade_b200.topk(rep, k, None)
    assert bool((counts == k).all())
    assert bool((vals[:, :-1] >= vals[:, 1:]).all())                                    # sorted by weight
    tie = vals[:, :-1] == vals[:, 1:]
    assert bool((ids[:, :-1][tie] < ids[:, 1:][tie]).all())                              # ties by ascending id
    assert torch.equal(torch.gather(rep, 1, ids.long()), vals)                           # weights are the row's
    pruned = torch.zeros_like(rep).scatter_(1, ids.long(), vals)
    ids2, vals2, _ = splade_b200.topk(pruned, k, None)                                   # idempotent
    assert torch.equal(ids2, ids) and torch.equal(vals2, vals)
    kth = vals[:, -1:]
    assert bool(((rep > kth).sum(1) < k).all())                                          # nothing larger left out
    rows = [0, 17, 511]
    _compare(ids[rows], vals[rows], counts[rows], O.topk_emit(rep[rows].cpu().numpy(), k, None), k)


def test_rows_larger_than_shared_memory_use_the_streaming_path(cuda_device):
    """V = 250 002 (the older XLM-R vocabulary size) does not fit one SM's shared memory: same results through the
    multi-pass global-memory variant."""
    import splade_b200
    g = torch.Generator().manual_seed(5)
    B, V, k = 3, 250002, 100
    rep = (torch.rand(B, V, generator=g) * 32).floor() / 32 * (torch.rand(B, V, generator=g) < 0.01)
    ids, vals, counts = splade_b200.topk(rep.to(cuda_device), k, None)
    _compare(ids, vals, counts, O.topk_emit(rep.numpy(), k, None), k)


def test_pathological_rows_take_the_exact_fallback(cuda_device):
    """Constant rows / huge tie sets overflow the fast path's candidate list and must be re-done by the exact
    kernel with identical semantics (lowest ids win the ties)."""
    import splade_b200
    V, k = 50000, 64
    rep = torch.zeros(5, V)
    rep[0] = 0.5                                   # every entry ties
    rep[1, ::2] = 0.25                             # 25 000 ties
    rep[2] = torch.rand(V, generator=torch.Generator().manual_seed(1))   # ordinary dense row in between
    rep[3, 100:5000] = 1.0                         # more ties at the k-th value than the sort buffer holds
    rep[3, :10] = 2.0
    rep[4, 7] = 3.0                                # almost empty row
    block = torch.zeros(V, dtype=torch.uint8)
    block[0] = 1
    ids, vals, counts = splade_b200.topk(rep.to(cuda_device), k, block.to(cuda_device))
    _compare(ids, vals, counts, O.topk_emit(rep.numpy(), k, block.numpy()), k)
    assert ids[0].tolist() == list(range(1, 65)) and ids[3, :9].tolist() == list(range(1, 10))
