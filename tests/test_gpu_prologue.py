"""-m gpu: the fused head prologue - transformers' ModernBertPredictionHead
``norm(act(dense(x)))`` (modeling_modernbert.py:493-502), called by the reference at src/model/splade_modern.py:69-73 -
forward and backward against the stock torch ops under autocast(bf16), at kernel level and through the drop-in
model (gradients of head.dense.weight / head.norm.weight / the trunk)."""
import copy

import pytest
import torch

from oracle import torch_reference as TR
from tests.gpu_util import assert_weights_parity
from tests.model_util import tiny_modernbert_dir

pytestmark = pytest.mark.gpu


def _torch_chain(x, dense_w, norm_w, eps):
    """The reference chain under autocast(bf16): bf16 Linear, bf16 GELU, fp32 LayerNorm (oracle/torch_reference.py)."""
    return TR.head_prologue_chain(x, dense_w, norm_w, eps, autocast=True)


@pytest.mark.parametrize("T,H", [(1000, 768), (64, 256), (16384, 768), (129, 1024)])
def test_prologue_forward_matches_autocast_chain(cuda_device, T, H):
    from splade_b200 import ops
    g = torch.Generator().manual_seed(T + H)
    x = torch.randn(T, H, generator=g).to(cuda_device)
    dense_w = (torch.randn(H, H, generator=g) * H ** -0.5).to(cuda_device)
    norm_w = (1.0 + 0.1 * torch.randn(H, generator=g)).to(cuda_device)
    eps = 1e-5
    rows_cap = max(T, 128)
    x16 = torch.zeros(rows_cap, H, dtype=torch.bfloat16, device=cuda_device)
    x16[:T] = x.to(torch.bfloat16)
    y, hidden, stats = ops.head_prologue_forward(x16, dense_w.to(torch.bfloat16).contiguous(), norm_w, eps, T)
    ry, rg, rh = _torch_chain(x, dense_w, norm_w, eps)
    assert rh.dtype == torch.float32 and ry.dtype == torch.bfloat16
    # dense output: the same bf16 value except where a different fp32 accumulation order crosses a rounding boundary
    dy = (y[:T].float() - ry.float()).abs()
    assert float((dy > 0).float().mean()) <= 2e-2
    assert float((dy - 2.0 ** -7 * ry.float().abs()).max()) <= 1e-6          # never more than one bf16 ulp
    # head input: bf16(LayerNorm); tolerance = one bf16 ulp of a value of magnitude <= 8, on nearly all cells
    dh = (hidden[:T].float() - rh.to(torch.bfloat16).float()).abs()
    assert float((dh > 2.0 ** -7 * rh.abs() + 1e-3).float().mean()) <= 1e-3
    torch.testing.assert_close(hidden[:T].float(), rh, rtol=2e-2, atol=2e-2)
    # statistics of the rows whose bf16 GELU values are identical
    rg32 = rg.float()
    torch.testing.assert_close(stats[:, 0], rg32.mean(-1), rtol=1e-2, atol=2e-3)
    torch.testing.assert_close(stats[:, 1], (rg32.var(-1, unbiased=False) + eps).rsqrt(), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("T,H", [(777, 768), (96, 256)])
def test_prologue_backward_matches_autocast_autograd(cuda_device, T, H):
    from splade_b200 import ops
    g = torch.Generator().manual_seed(T * 3 + H)
    x = torch.randn(T, H, generator=g).to(cuda_device).requires_grad_(True)
    dense_w = (torch.randn(H, H, generator=g) * H ** -0.5).to(cuda_device).requires_grad_(True)
    norm_w = (1.0 + 0.1 * torch.randn(H, generator=g)).to(cuda_device).requires_grad_(True)
    d_hidden = torch.randn(T, H, generator=g).to(cuda_device)
    eps = 1e-5
    # exact (fp32 arithmetic) gradients and the reference's own autocast gradients
    h32 = TR.head_prologue_chain(x, dense_w, norm_w, eps, autocast=False)[2]
    ex = torch.autograd.grad(h32, (x, dense_w, norm_w), d_hidden)
    _, _, rh = _torch_chain(x, dense_w, norm_w, eps)
    ac = torch.autograd.grad(rh, (x, dense_w, norm_w), d_hidden)

    rows_cap = max(T, 128)
    x16 = torch.zeros(rows_cap, H, dtype=torch.bfloat16, device=cuda_device)
    x16[:T] = x.detach().to(torch.bfloat16)
    w16 = dense_w.detach().to(torch.bfloat16).contiguous()
    y, hidden, stats = ops.head_prologue_forward(x16, w16, norm_w.detach(), eps, T)
    dy, dx, d_norm = ops.head_prologue_backward(d_hidden, y, stats, norm_w.detach(), w16.t().contiguous(), T)
    d_dense = torch.mm(dy[:T].float().t(), x16[:T].float())

    def rel(a, b):
        return float((a.float() - b.float()).norm() / b.float().norm())

    # within the north-star tolerance of the reference's autocast gradients, and no further from the exact gradients
    # than the reference's own autocast backward is
    for name, got, a, e in (("dx", dx[:T], ac[0], ex[0]), ("d_dense", d_dense, ac[1], ex[1]),
                            ("d_norm", d_norm, ac[2], ex[2])):
        assert rel(got, a) <= 1e-2, (name, rel(got, a))
        assert rel(got, e) <= max(1.5 * rel(a, e), 2e-3), (name, rel(got, e), rel(a, e))
    # run-to-run determinism of the two-stage dgamma sum
    _, _, d_norm2 = ops.head_prologue_backward(d_hidden, y, stats, norm_w.detach(), None, T)
    assert torch.equal(d_norm, d_norm2)


def test_prologue_argument_errors(cuda_device):
    from splade_b200 import ops
    x16 = torch.zeros(128, 320, dtype=torch.bfloat16, device=cuda_device)
    with pytest.raises(RuntimeError, match="hidden size"):
        ops.head_prologue_forward(x16, torch.zeros(320, 320, dtype=torch.bfloat16, device=cuda_device),
                                  torch.ones(320, device=cuda_device), 1e-5, 100)


@pytest.fixture(scope="module")
def models256(cuda_device):
    import splade_b200
    d = tiny_modernbert_dir(vocab_size=5000, hidden=256, heads=4, bias_offset=-0.3)
    ours = splade_b200.SPLADEModernBERT(model_name=d).to(cuda_device)
    ref = TR.ReferenceSpladeModule(copy.deepcopy(ours.model)).to(cuda_device)
    return ours, ref


def _ids(B, S, seed, dev, V=5000):
    from splade_b200.training import synthetic_triplet_batch
    b = synthetic_triplet_batch(B, S, S, 1, V, V - 1, seed)
    return b["positive_input_ids"].to(dev), b["positive_attention_mask"].to(dev)


def test_model_uses_the_fused_prologue_and_matches(models256, cuda_device, monkeypatch):
    from splade_b200 import ops
    ours, ref = models256
    assert ops.prologue_of(ours.model.head) is not None
    ours.eval(), ref.eval()
    ids, mask = _ids(8, 64, 1, cuda_device)
    calls = []
    real = ops.head_prologue_forward
    monkeypatch.setattr(ops, "head_prologue_forward", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        rep, tokw = ours(ids, mask)
        r_rep, r_tokw = ref(ids, mask)
    assert calls, "the fused prologue did not run under autocast(bf16)"
    # a bf16 ulp in the head input moves many logits a little: same tolerance as the north star (1e-2), plus the
    # active sets must agree except at the threshold
    torch.testing.assert_close(rep, r_rep, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(tokw + 0, r_tokw, rtol=1e-2, atol=1e-2)
    disagree = ((rep > 0) != (r_rep > 0)).float().mean()
    assert float(disagree) <= 1e-3
    # outside autocast the transformers module runs (fp32 evaluation path of the reference)
    calls.clear()
    with torch.no_grad():
        ours(ids, mask)
    assert not calls
    # and the switch
    monkeypatch.setenv("SPLADE_B200_HEAD_PROLOGUE", "0")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        rep_off, _ = ours(ids, mask)
    assert not calls
    assert_weights_parity(rep_off.cpu().numpy(), r_rep.cpu().numpy(), "prologue off = module path")


def test_model_gradients_through_the_fused_prologue(models256, cuda_device):
    """head.dense.weight / head.norm.weight (and everything upstream) at 1e-2... against the
    reference's autocast backward, whose own bf16 rounding noise is ~1 % (see test_gpu_model.py): 2e-2 like there."""
    import splade_b200
    ours, ref = models256
    ours.train(), ref.train()
    ref.load_state_dict(ours.state_dict())
    q_ids, q_mask = _ids(8, 16, 2, cuda_device)
    p_ids, p_mask = _ids(8, 64, 3, cuda_device)
    n_ids, n_mask = _ids(8, 64, 4, cuda_device)
    loss_fn = splade_b200.SPLADELossV33()
    step = 5000
    lam = [loss_fn._lambda_schedule(step, x) for x in (loss_fn.lambda_q, loss_fn.lambda_d, loss_fn.lambda_neg)]
    for m in (ours, ref):
        m.zero_grad()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a, _ = ours(q_ids, q_mask)
        p, _ = ours(p_ids, p_mask)
        n, _ = ours(n_ids, n_mask)
        loss, _ = loss_fn(a, p, n, global_step=step)
        ra, _ = ref(q_ids, q_mask)
        rp, _ = ref(p_ids, p_mask)
        rn, _ = ref(n_ids, n_mask)
        r_loss = TR.v33_loss(ra, rp, rn, *lam)
    loss.backward()
    r_loss.backward()
    assert float(loss) == pytest.approx(float(r_loss), rel=1e-2)
    grads = dict(ours.named_parameters())
    for name, rp_ in ref.named_parameters():
        g, rg = grads[name].grad, rp_.grad
        assert g is not None, name
        scale = float(rg.abs().max())
        err = float((g - rg).abs().max())
        assert err <= 2e-2 * scale + 1e-9, (name, err, scale)
        rel = float((g - rg).norm() / rg.norm().clamp_min(1e-12))
        assert rel <= 2e-2, (name, rel)
    for name in ("model.head.dense.weight", "model.head.norm.weight"):
        assert float(grads[name].grad.abs().max()) > 0


def test_prologue_forward_against_the_numpy_oracle(cuda_device):
    """The CUDA prologue against oracle/splade_oracle.py::head_prologue (autocast form), itself pinned to the
    transformers module in the CPU suite."""
    import numpy as np

    from oracle import splade_oracle as O
    from splade_b200 import ops
    g = torch.Generator().manual_seed(11)
    T, H = 200, 256
    x = torch.randn(T, H, generator=g)
    dense_w = torch.randn(H, H, generator=g) * H ** -0.5
    norm_w = 1.0 + 0.1 * torch.randn(H, generator=g)
    ry, rg, rh = O.head_prologue(x.numpy(), dense_w.numpy(), norm_w.numpy(), 1e-5, mode="bf16")
    x16 = x.to(torch.bfloat16).to(cuda_device).contiguous()
    y, hidden, stats = ops.head_prologue_forward(x16, dense_w.to(torch.bfloat16).to(cuda_device).contiguous(),
                                                 norm_w.to(cuda_device), 1e-5, T)
    y = y[:T].float().cpu().numpy()
    assert float(np.mean(y != ry)) <= 2e-2                                   # accumulation order at rounding boundaries
    assert float(np.max(np.abs(y - ry) - 2.0 ** -7 * np.abs(ry))) <= 1e-6    # never more than one bf16 ulp
    rows_same = (y == ry).all(axis=1)
    assert rows_same.sum() >= 1
    want = O.bf16_round(rh)[rows_same]
    got = hidden[:T].float().cpu().numpy()[rows_same]
    bad = np.abs(got - want) > 2.0 ** -7 * np.abs(want) + 1e-3              # one bf16 ulp
    assert float(bad.mean()) <= 1e-3
    np.testing.assert_allclose(stats[:, 0].cpu().numpy()[rows_same], rg[rows_same].mean(axis=1), rtol=1e-3, atol=1e-4)
