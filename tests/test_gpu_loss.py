"""-m gpu: SPLADELossV33 drop-in (C-ABI loss kernels) vs golden vectors from the reference and vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import splade_oracle as O
from oracle import torch_reference as TR

pytestmark = pytest.mark.gpu

LD_KEYS = ("infonce", "flops_q", "flops_d", "flops_neg", "lambda_q", "lambda_d", "lambda_neg", "kd",
           "margin_mse", "nonzero_q", "nonzero_d")


def _run(loss_fn, q, p, n, step, dev, **kw):
    qt, pt, nt = (torch.from_numpy(np.asarray(a, np.float32)).to(dev).requires_grad_(True) for a in (q, p, n))
    kw = {k: torch.from_numpy(np.asarray(v, np.float32)).to(dev) for k, v in kw.items()}
    loss, d = loss_fn(qt, pt, nt, global_step=step, **kw)
    loss.backward()
    return float(loss), d, (qt.grad.cpu().numpy(), pt.grad.cpu().numpy(), nt.grad.cpu().numpy())


def test_g2_known_answers(cuda_device, golden):
    import splade_b200
    g = golden("g2_loss")
    for step, total in ((0, 0.791450977), (10000, 0.797582209), (20000, 0.815975964)):
        loss, d, grads = _run(splade_b200.SPLADELossV33(), g["q"], g["p"], g["n"], step, cuda_device)
        assert loss == pytest.approx(total, rel=1e-5)
        assert set(d) == set(LD_KEYS) and all(isinstance(v, float) for v in d.values())
        for key in LD_KEYS:
            assert d[key] == pytest.approx(float(g[f"ld_{key}_{step}"]), rel=1e-5, abs=1e-7), key
        for got, key in zip(grads, ("dq", "dp", "dn")):
            np.testing.assert_allclose(got, g[f"{key}_{step}"], rtol=1e-4, atol=1e-6)


def test_g3_multineg_margin_mse(cuda_device, golden):
    import splade_b200
    g = golden("g3_multineg")
    fn = splade_b200.SPLADELossV33(lambda_margin_mse=0.5, lambda_neg=0.01, flops_warmup_steps=0)
    loss, d, grads = _run(fn, g["q"], g["p"], g["n"], 5, cuda_device, teacher_pos_scores=g["tp"],
                          teacher_neg_scores=g["tn"])
    assert loss == pytest.approx(2.189625740, rel=1e-5)
    assert d["margin_mse"] == pytest.approx(2.38499999, rel=1e-5)
    for got, key in zip(grads, ("dq", "dp", "dn")):
        np.testing.assert_allclose(got, g[key], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(fn.get_avg_nonzero(), g["avg_nonzero"], rtol=1e-6)


@pytest.mark.parametrize("name", ["loss_rand_k1", "loss_rand_k3", "loss_rand_kd", "loss_rand_mmse"])
def test_random_cases_vs_reference(cuda_device, golden, name):
    import splade_b200
    g = golden(name)
    ctor = {k[len("ctor_"):]: float(v) for k, v in g.items() if k.startswith("ctor_")}
    kw = {k: g[k] for k in ("teacher_scores", "teacher_pos_scores", "teacher_neg_scores") if k in g}
    loss, d, grads = _run(splade_b200.SPLADELossV33(**ctor), g["q"], g["p"], g["n"], int(g["step"]), cuda_device,
                          **kw)
    assert loss == pytest.approx(float(g["loss"]), rel=2e-5)
    for key in LD_KEYS:
        assert d[key] == pytest.approx(float(g["ld_" + key]), rel=2e-5, abs=1e-6), key
    for got, key in zip(grads, ("dq", "dp", "dn")):
        np.testing.assert_allclose(got, g[key], rtol=5e-4, atol=1e-6)


def _reps(n, V, seed, density, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, V, generator=g) * 1.4
    return (x * (torch.rand(n, V, generator=g) < density)).to(dev)


@pytest.mark.parametrize("k", [1, 7])
def test_training_shape_fp32_and_autocast(cuda_device, k):
    """BASELINE config 4 loss shape ([64, 50 000], k hard negatives): fp32 mode vs the float64 oracle, and
    autocast mode vs the torch statement of the reference under autocast on the same device."""
    import splade_b200
    Bq, V = 64, 50000
    q = _reps(Bq, V, 1, 0.05, cuda_device)
    p = _reps(Bq, V, 2, 0.08, cuda_device)
    n = _reps(Bq * k, V, 3, 0.08, cuda_device)
    n = n.view(Bq, k, V) if k > 1 else n
    fn = splade_b200.SPLADELossV33()
    step = 10000
    lq, ld, ln = (fn._lambda_schedule(step, x) for x in (fn.lambda_q, fn.lambda_d, fn.lambda_neg))

    # fp32
    loss, d, grads = _run(fn, q.cpu().numpy(), p.cpu().numpy(), n.cpu().numpy(), step, cuda_device)
    ref_loss, ref_d, ref_g = O.v33_loss(q.cpu().numpy(), p.cpu().numpy(), n.cpu().numpy(), global_step=step)
    assert loss == pytest.approx(ref_loss, rel=1e-4)
    for key in LD_KEYS:
        assert d[key] == pytest.approx(ref_d[key], rel=1e-4, abs=1e-6), key
    for got, ref in zip(grads, ref_g):
        np.testing.assert_allclose(got, ref, rtol=1e-3, atol=1e-7 + 1e-4 * np.abs(ref).max())

    # autocast: the reference's in-batch scores are bf16 -> compare with the same chain on this GPU
    outs = []
    for impl in ("ours", "ref"):
        qt, pt, nt = (t.clone().requires_grad_(True) for t in (q, p, n))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if impl == "ours":
                l, _ = fn(qt, pt, nt, global_step=step)
            else:
                l = TR.v33_loss(qt, pt, nt, lq, ld, ln, fn.temperature)
        l.backward()
        outs.append((l.detach(), qt.grad, pt.grad, nt.grad))
    assert float(outs[0][0]) == pytest.approx(float(outs[1][0]), rel=1e-2)   # north_star bf16 tolerance
    for a, r, name in zip(outs[0][1:], outs[1][1:], ("dq", "dp", "dn")):
        scale = float(r.abs().max())
        assert float((a - r).abs().max()) <= 1e-2 * scale + 1e-8, name


def test_gathered_negatives_virtual_world(cuda_device):
    """Cross-GPU negatives, emulated in one process: rank r scores its queries against the rank-order
    concatenation of all positives with labels r*B + i.  The mean over ranks equals the reference's InfoNCE on
    the concatenated batch (SURVEY.md §8(e) oracle), and the summed document gradients match."""
    import splade_b200
    from splade_b200.ops import V33LossFunction
    W_, Bq, V = 4, 8, 3000
    q_all = _reps(W_ * Bq, V, 11, 0.1, cuda_device)
    p_all = _reps(W_ * Bq, V, 12, 0.1, cuda_device)
    n_all = _reps(W_ * Bq, V, 13, 0.1, cuda_device)
    cfg = dict(round_bf16=False, temperature=1.0, kd_temperature=1.0, lambda_q=0.0, lambda_d=0.0, lambda_neg=0.0,
               lambda_kd=0.0, lambda_mmse=0.0)
    losses, dp_sum = [], torch.zeros_like(p_all)
    for r in range(W_):
        sl = slice(r * Bq, (r + 1) * Bq)
        docs = p_all.clone().requires_grad_(True)
        loss, _ = V33LossFunction.apply(q_all[sl], docs, n_all[sl].unsqueeze(1), None, None, None,
                                        dict(cfg, label_offset=r * Bq))
        loss.backward()
        losses.append(float(loss))
        dp_sum += docs.grad
    # reference on the concatenated batch: in-batch part over all W*B docs, each query keeps its own hard negative
    ref_loss, _, ref_g = O.v33_loss(q_all.cpu().numpy(), p_all.cpu().numpy(), n_all.cpu().numpy(), lambda_q=0.0,
                                    lambda_d=0.0)
    assert np.mean(losses) == pytest.approx(ref_loss, rel=1e-5)
    np.testing.assert_allclose((dp_sum / W_).cpu().numpy(), ref_g[1], rtol=1e-3, atol=1e-7)


def test_loss_dict_is_one_sync_and_ema(cuda_device):
    import splade_b200
    fn = splade_b200.SPLADELossV33()
    q, p, n = (_reps(4, 512, s, 0.2, cuda_device) for s in (1, 2, 3))
    for step in range(3):
        _, d = fn(q, p, n, global_step=step)
    nq = float((q > 0).sum(1).float().mean())
    want = 0.0
    for _ in range(3):
        want = 0.9 * want + 0.1 * nq
    assert fn.get_avg_nonzero()[0] == pytest.approx(want, rel=1e-6)
    assert d["nonzero_q"] == pytest.approx(nq)


def test_inbatch_scores_bit_exact_at_training_shape(cuda_device):
    """torch.mm under autocast (bf16 in, fp32 accumulate, bf16 out) vs splade_scores with round_bf16 at
    [64, 50 000] x [50 000, 64]: identical bf16 bit patterns (up to rounding-boundary cases)."""
    import splade_b200
    q = _reps(64, 50000, 5, 0.05, cuda_device)
    p = _reps(64, 50000, 6, 0.08, cuda_device)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref = torch.mm(q, p.t())
    ours = splade_b200.scores(q, p, True).to(torch.bfloat16)
    assert ref.dtype == torch.bfloat16
    assert int((ours != ref).sum()) <= 2
    exact = q.double() @ p.double().t()
    torch.testing.assert_close(splade_b200.scores(q, p, False).double(), exact, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("Bq,N,V", [(5, 7, 130), (64, 64, 50000), (130, 300, 1001), (64, 512, 50000), (1, 1, 8)])
def test_tensor_core_scores_shapes(cuda_device, Bq, N, V):
    """round_bf16 path = tcgen05 split-K GEMM (scores_tc.cu): bf16-rounded operands, fp32 accumulation, against a
    float64 product of the same rounded operands; covers ragged M/N tiles, V not a multiple of 8 or 64."""
    import splade_b200
    q = _reps(Bq, V, 31, 0.3, cuda_device)
    d = _reps(N, V, 32, 0.3, cuda_device)
    raw = splade_b200.scores(q, d, True)
    want = q.to(torch.bfloat16).double() @ d.to(torch.bfloat16).double().t()
    torch.testing.assert_close(raw.double(), want, rtol=2e-5, atol=1e-4)
    raw2 = splade_b200.scores(q, d, True)
    assert torch.equal(raw, raw2)          # fixed-order split-K reduction: run-to-run deterministic


@pytest.mark.parametrize("Bq,N,k,off,V", [(64, 64, 1, 0, 50000), (64, 512, 1, 128, 50000), (16, 16, 7, 0, 50000),
                                          (24, 200, 3, 40, 4100), (5, 9, 2, 3, 260)])
def test_tensor_core_backward_matches_the_fp32_backward(cuda_device, monkeypatch, Bq, N, k, off, V):
    """Autocast mode: dq / ddocs / dnegs from splade_loss_bwd_tc (tcgen05 on the bf16 operands the forward packed,
    G rounded to bf16 - what the reference's autocast backward of torch.mm computes, minus its bf16 output rounding)
    against splade_loss_bwd (fp32 CUDA cores) on the same forward.  Stated tolerance: 1e-2 of the gradient scale
    (north_star); measured ~2e-3.  Covers gathered documents (N > Bq, label offset), N and Bq that are not multiples
    of 64 / 128, several hard negatives, a vocabulary tail."""
    from splade_b200.ops import V33LossFunction
    q = _reps(Bq, V, 21, 0.05, cuda_device)
    docs = _reps(N, V, 22, 0.08, cuda_device)
    negs = _reps(Bq * k, V, 23, 0.08, cuda_device).view(Bq, k, V)
    cfg = dict(round_bf16=True, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3, lambda_neg=3e-3,
               lambda_kd=0.0, lambda_mmse=0.0, label_offset=off)
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("SPLADE_B200_LOSS_BWD_TC", tc)
        a, b, c = (t.clone().requires_grad_(True) for t in (q, docs, negs))
        loss, _ = V33LossFunction.apply(a, b, c, None, None, None, cfg)
        (loss * 1.7).backward()
        outs[tc] = (float(loss), a.grad, b.grad, c.grad)
    assert outs["1"][0] == outs["0"][0]                       # same forward
    for got, ref, name in zip(outs["1"][1:], outs["0"][1:], ("dq", "ddocs", "dnegs")):
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        assert err <= 1e-2 * scale + 1e-9, (name, err, scale)
        rel = float((got - ref).norm() / ref.norm().clamp_min(1e-20))
        assert rel <= 5e-3, (name, rel)
    torch.testing.assert_close(outs["1"][3], outs["0"][3], rtol=1e-6, atol=1e-12)   # dnegs has no GEMM term
