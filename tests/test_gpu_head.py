"""-m gpu: fused head forward/backward through the C-ABI vs the oracle, golden vectors and the torch
autocast chain on the same device."""
import numpy as np
import pytest
import torch

from oracle import splade_oracle as O
from oracle import torch_reference as TR
from tests.gpu_util import (assert_argmax_parity, assert_weights_parity, make_head_inputs, query_like_mask,
                            run_head)

pytestmark = pytest.mark.gpu


def _check_forward(hidden, W, bias, mask, want_tokw=True, what="", pack_mode=None):
    rep, arg, tokw = run_head(hidden, W, bias, mask, want_tokw, pack_mode=pack_mode)
    torch.cuda.synchronize()
    h, w, b, m = (t.cpu().numpy() for t in (hidden, W, bias, mask))
    ref_rep, ref_arg, ref_tokw = O.head_forward(h, w, b, m, mode="bf16")
    near = O.near_tie_mask(h, w, b, m, mode="bf16")
    flips = assert_argmax_parity(arg.cpu().numpy(), ref_arg, near, what)
    # weights: identical bf16 logit -> log1p agrees to an fp32 ulp; near-tie flips may move by one bf16 ulp
    got = rep.cpu().numpy()
    assert_weights_parity(got, ref_rep, what)
    if want_tokw:
        assert_weights_parity(tokw.cpu().numpy(), ref_tokw, what + " token_weights", max_flip_frac=1e-3)
    else:
        assert tokw is None
    return flips


@pytest.mark.parametrize("name", ["head_bf16", "head_sparse", "head_s300"])
def test_golden_forward_backward(cuda_device, golden, name):
    g = golden(name)
    t = {k: torch.from_numpy(v).to(cuda_device) for k, v in g.items()}
    rep, arg, tokw = run_head(t["hidden"], t["W"], t["bias"], t["mask"])
    near = O.near_tie_mask(g["hidden"], g["W"], g["bias"], g["mask"], mode="bf16")
    assert_argmax_parity(arg.cpu().numpy(), g["argmax"], near, name)
    assert_weights_parity(rep.cpu().numpy(), g["sparse_repr"], name)
    assert_weights_parity(tokw.cpu().numpy(), g["token_weights"], name + " token_weights", max_flip_frac=1e-3)

    import splade_b200
    # feed the REFERENCE's rep/argmax so that the backward check is independent of near-tie flips
    d_hidden, dW, dbias = splade_b200.head_backward(
        t["grad_out"], t["sparse_repr"].contiguous(), t["argmax"].contiguous(),
        t["hidden"].to(torch.bfloat16).contiguous(), t["W"].to(torch.bfloat16).contiguous())
    np.testing.assert_allclose(d_hidden.cpu().numpy(), g["d_hidden"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dW.cpu().numpy(), g["dW"], rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dbias.cpu().numpy(), g["dbias"], rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize("B,S,H,V,kw", [
    (1, 1, 64, 1, {}),                       # degenerate
    (3, 7, 64, 130, {}),                     # many tiny docs per tile, vocab tail
    (8, 64, 128, 50000, {}),                 # BASELINE config 1 head shape
    (5, 100, 64, 777, {"holes": True}),      # S not a multiple of 16/32, non-suffix mask, 2 docs per tile
    (9, 250, 192, 1000, {}),                 # window clamp at the end of a tile
    (2, 300, 64, 260, {}),                   # document spans two token tiles (atomic path)
    (2, 700, 128, 200, {"bias_offset": -1.0}),
    (4, 256, 768, 2048, {"bias_offset": -2.0}),  # sparse regime at the real hidden size
])
def test_forward_vs_oracle(cuda_device, B, S, H, V, kw):
    inputs = make_head_inputs(B, S, H, V, seed=B * 1000 + S, device=cuda_device, **kw)
    _check_forward(*inputs, what=f"B{B} S{S} H{H} V{V}")
    _check_forward(*inputs, want_tokw=False, what=f"B{B} S{S} H{H} V{V} (no tokw)")


def test_all_padding_document_and_zero_rows(cuda_device):
    hidden, W, bias, mask = make_head_inputs(4, 32, 64, 300, seed=5, device=cuda_device)
    mask[2] = 0  # a document with no real token at all
    rep, arg, tokw = run_head(hidden, W, bias, mask)
    assert float(rep[2].abs().max()) == 0.0 and int(arg[2].abs().max()) == 0 and float(tokw[2].abs().max()) == 0.0
    _check_forward(hidden, W, bias, mask, what="all-padding doc")


def test_headline_shape_vs_torch_autocast(cuda_device):
    """BASELINE config 2 (B=64, S=256, H=768, V=50 000) against the reference chain under autocast on the same
    GPU (cuBLAS bf16 GEMM + ATen): bit-exact argmax away from near-ties, weights within fp32 ulps."""
    B, S, H, V = 64, 256, 768, 50000
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=cuda_device)
    rep, arg, tokw = run_head(hidden, W, bias, mask)
    ref_rep, ref_arg, ref_tokw = TR.head_chain_autocast(hidden, W, bias, mask)
    # near-ties from the torch chain itself: top-2 over the sequence of relu(bf16 logits)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = torch.nn.functional.linear(hidden, W, bias)
    x = torch.relu(logits).float() * mask.unsqueeze(-1).float()
    del logits
    top2 = torch.topk(x, 2, dim=1).values
    del x
    near = (top2[:, 0] - top2[:, 1]) <= top2[:, 0].abs().clamp_min(2.0 ** -126) * 2.0 ** -7
    flips = assert_argmax_parity(arg.cpu().numpy(), ref_arg.cpu().numpy(), near.cpu().numpy(), "C2 vs autocast")
    wflips = assert_weights_parity(rep.cpu().numpy(), ref_rep.cpu().numpy(), "C2 vs autocast")
    assert_weights_parity(tokw.cpu().numpy(), ref_tokw.cpu().numpy(), "C2 token_weights", max_flip_frac=1e-3)
    print(f"C2 vs cuBLAS+ATen autocast chain: argmax flips {flips} of {B * V} (all at near-ties, "
          f"{int(near.sum())} near-tie cells); weights off by one bf16 ulp: {wflips}")


def test_size_independent_properties(cuda_device):
    """Full-size properties: a document's row does not depend on its neighbours in the batch, on the order of
    the batch, or on what sits at padded positions; doubling every logit's input keeps the argmax."""
    B, S, H, V = 64, 256, 768, 50000
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=3, device=cuda_device)
    rep, arg, _ = run_head(hidden, W, bias, mask, want_tokw=False)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).to(cuda_device)
    rep_p, arg_p, _ = run_head(hidden[perm].contiguous(), W, bias, mask[perm].contiguous(), want_tokw=False)
    assert torch.equal(rep_p, rep[perm]) and torch.equal(arg_p, arg[perm])
    sub = slice(10, 13)
    rep_s, arg_s, _ = run_head(hidden[sub].contiguous(), W, bias, mask[sub].contiguous(), want_tokw=False)
    assert torch.equal(rep_s, rep[sub]) and torch.equal(arg_s, arg[sub])
    junk = hidden.clone()
    junk[mask == 0] = 37.0
    rep_j, arg_j, _ = run_head(junk, W, bias, mask, want_tokw=False)
    assert torch.equal(rep_j, rep) and torch.equal(arg_j, arg)
    assert float(rep.min()) >= 0.0 and int(arg.min()) >= 0
    lengths = mask.sum(1)
    assert bool((arg < lengths[:, None].clamp_min(1)).all())


def test_backward_vs_oracle_random(cuda_device):
    import splade_b200
    B, S, H, V = 6, 48, 128, 999
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=9, device=cuda_device, bias_offset=-0.5)
    rep, arg, _ = run_head(hidden, W, bias, mask, want_tokw=False)
    g = torch.randn(B, V, generator=torch.Generator().manual_seed(1)).to(cuda_device)
    d_hidden, dW, dbias = splade_b200.head_backward(g, rep, arg, hidden.to(torch.bfloat16).contiguous(),
                                                    W.to(torch.bfloat16).contiguous())
    h, w, b, m = (t.cpu().numpy() for t in (hidden, W, bias, mask))
    rdh, rdW, rdb = O.head_backward(g.cpu().numpy(), h, w, b, m, mode="bf16")
    near = O.near_tie_mask(h, w, b, m, mode="bf16")
    if not near.any():  # identical routing -> tight check
        np.testing.assert_allclose(d_hidden.cpu().numpy(), rdh, rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(dW.cpu().numpy(), rdW, rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(dbias.cpu().numpy(), rdb, rtol=1e-2, atol=1e-4)


def test_autograd_function_matches_torch_chain(cuda_device):
    """End to end through torch.autograd: gradients of hidden / weight / bias against the reference chain under
    autocast on the same device (stated bf16 tolerance: <= 1e-2 relative to the gradient scale)."""
    import splade_b200
    B, S, H, V = 8, 64, 128, 5000
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=4, device=cuda_device)
    gout = torch.randn(B, V, generator=torch.Generator().manual_seed(2)).to(cuda_device)
    outs = []
    for impl in ("ours", "ref"):
        h = hidden.clone().requires_grad_(True)
        w = W.clone().requires_grad_(True)
        b = bias.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if impl == "ours":
                rep, tokw = splade_b200.splade_head(h, w, b, mask)
            else:
                rep, _, tokw = TR.head_chain(h, w, b, mask)
        (rep * gout).sum().backward()
        outs.append((rep.detach(), tokw.detach(), h.grad, w.grad, b.grad))
    for a, r, name in zip(outs[0], outs[1], ("rep", "tokw", "d_hidden", "dW", "dbias")):
        assert a.dtype == r.dtype and a.shape == r.shape, name
        scale = float(r.abs().max())
        err = float((a - r).abs().max())
        assert err <= 1e-2 * scale + 1e-6, (name, err, scale)


# ---------------------------------------------------------------------------------------------------------
# one backward, one dW: several head calls sharing the decoder matrix (splade_head_bwd_multi / _dhidden / _dw)
# ---------------------------------------------------------------------------------------------------------
def _segments(device, shapes, H, V, bias_offset, seed=11):
    """[(g_rep, rep, argmax, hidden_bf16)] for head calls of different (B, S) on ONE (W, bias)."""
    _, W, bias, _ = make_head_inputs(1, 8, H, V, seed=seed, device=device, bias_offset=bias_offset)
    w16 = W.to(torch.bfloat16).contiguous()
    segs = []
    for i, (B, S) in enumerate(shapes):
        if B == 0:  # an empty head call (e.g. a rank whose shard ran dry)
            segs.append((torch.empty(0, V, device=device), torch.empty(0, V, device=device),
                         torch.empty(0, V, dtype=torch.int32, device=device),
                         torch.empty(0, S, H, dtype=torch.bfloat16, device=device)))
            continue
        hidden, _, _, mask = make_head_inputs(B, S, H, V, seed=seed + 1 + i, device=device)
        rep, arg, _ = run_head(hidden, W, bias, mask, want_tokw=False)
        g = torch.randn(B, V, generator=torch.Generator().manual_seed(seed + 100 + i)).to(device)
        segs.append((g, rep, arg, hidden.to(torch.bfloat16).contiguous()))
    return segs, w16


@pytest.mark.parametrize("shapes,H,V,off", [
    ([(4, 16), (4, 64), (8, 40)], 128, 1999, -0.5),      # q / pos / neg of different lengths, vocabulary tail
    ([(3, 7)], 64, 33, 0.0),                             # one segment = splade_head_bwd
    ([(2, 300), (5, 9), (1, 1), (0, 4)], 64, 1000, -1.0),  # long documents, a 1x1 call, an EMPTY segment
    ([(150, 8), (150, 8)], 64, 520, -0.3),               # 300 documents: two staging passes of the dW sweep
    ([(2, 64), (2, 256), (2, 256)], 768, 4096, -2.0),    # real hidden size, trained-like sparsity
])
def test_multi_segment_backward_equals_the_sum_of_single_calls(cuda_device, shapes, H, V, off):
    import splade_b200
    segs, w16 = _segments(cuda_device, shapes, H, V, off)
    singles = [splade_b200.head_backward(g, rep, arg, h, w16) for g, rep, arg, h in segs if h.shape[0] > 0]
    dhs, dW, dbias = splade_b200.head_backward_multi(segs, w16)
    torch.cuda.synchronize()
    j = 0
    for (g, rep, arg, h), dh in zip(segs, dhs):
        if h.shape[0] == 0:
            continue
        assert torch.equal(dh, singles[j][0]), "dHidden of a segment must not depend on the other segments"
        j += 1
    ref_dW = torch.stack([s[1].double() for s in singles]).sum(0)
    ref_db = torch.stack([s[2].double() for s in singles]).sum(0)
    scale = float(ref_dW.abs().max())
    assert float((dW.double() - ref_dW).abs().max()) <= 1e-5 * scale + 1e-7
    assert float((dbias.double() - ref_db).abs().max()) <= 1e-5 * float(ref_db.abs().max()) + 1e-7

    # two-phase form (what autograd drives): bit-identical to the three-launch form, also for a NULL dHidden
    live = [s for s in segs if s[3].shape[0] > 0]
    phase = [splade_b200.head_backward_dhidden(g, rep, arg, h, w16, want_d_hidden=(i != 0))
             for i, (g, rep, arg, h) in enumerate(live)]
    assert phase[0][0] is None
    for (dh2, _), dh in zip(phase[1:], [d for d, s in zip(dhs, segs) if s[3].shape[0] > 0][1:]):
        assert torch.equal(dh2, dh)
    dW2, db2 = splade_b200.head_backward_dw([(g, rep, arg, h, bm) for (g, rep, arg, h), (_, bm) in zip(live, phase)],
                                            H, V)
    assert torch.equal(dW2, dW) and torch.equal(db2, dbias)

    # the bitmap made right after the forward (what autograd passes) gives bit-identical results to the one the
    # library derives itself, in the three-launch and in the two-phase form
    act = [splade_b200.rep_bitmap(s[1]) for s in segs]
    dhs4, dW4, db4 = splade_b200.head_backward_multi([(*s, a) for s, a in zip(segs, act)], w16)
    assert torch.equal(dW4, dW) and torch.equal(db4, dbias) and all(torch.equal(a, b) for a, b in zip(dhs4, dhs))
    g0, rep0, arg0, h0 = live[-1]
    act0 = splade_b200.rep_bitmap(rep0)
    dh5, bm5 = splade_b200.head_backward_dhidden(g0, rep0, arg0, h0, w16, active=act0)
    assert bm5 is act0 and torch.equal(bm5, phase[-1][1]) and torch.equal(dh5, [d for d, s in zip(dhs, segs) if s[3].shape[0] > 0][-1])

    # the combined entry autograd's last node uses (its dHidden + the sweep over all segments, concurrently)
    last5 = (g0, rep0, arg0, h0, act0)
    all5 = [(*s_, a) for s_, a in zip(live, [splade_b200.rep_bitmap(s_[1]) for s_ in live])]
    dh6, dW6, db6 = splade_b200.ops.head_backward_last(last5, all5, w16)
    assert torch.equal(dh6, dh5) and torch.equal(dW6, dW) and torch.equal(db6, dbias)

    # accumulate: adds into the caller's gradient buffers (same sweep, starting from their contents)
    base_W = torch.randn(V, H, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    base_b = torch.randn(V, generator=torch.Generator().manual_seed(6)).to(cuda_device)
    acc_W, acc_b = base_W.clone(), base_b.clone()
    splade_b200.head_backward_multi(segs, w16, accumulate_into=(acc_W, acc_b))
    # (the sweep continues FROM the buffer's value, ((base + t1) + t2) + ..., so it differs from base + (t1 + t2 + ...)
    # by fp32 rounding at the magnitude of base ~ N(0,1), not of the gradient)
    torch.testing.assert_close(acc_W, base_W + dW, rtol=1e-5, atol=5e-5)
    torch.testing.assert_close(acc_b, base_b + dbias, rtol=1e-5, atol=5e-5)

    # deterministic: the same call again is bit-identical
    dhs3, dW3, db3 = splade_b200.head_backward_multi(segs, w16)
    assert torch.equal(dW3, dW) and torch.equal(db3, dbias) and all(torch.equal(a, b) for a, b in zip(dhs3, dhs))


def test_three_head_calls_one_backward_one_dw_sweep(cuda_device, monkeypatch):
    """The reference's micro-step pattern through autograd (train_v33_ddp.py:339-343, 364): three head calls on the
    tied matrix, one backward => ONE dW sweep over three segments; gradients equal the un-deferred path."""
    import splade_b200
    from splade_b200 import ops
    H, V = 128, 3000
    _, W, bias, _ = make_head_inputs(1, 8, H, V, seed=21, device=cuda_device, bias_offset=-0.5)
    ins = [make_head_inputs(B, S, H, V, seed=22 + i, device=cuda_device)[::3] for i, (B, S) in
           enumerate(((4, 16), (4, 48), (8, 48)))]
    sweeps = []
    real_last = ops.head_backward_last
    monkeypatch.setattr(ops, "head_backward_last",
                        lambda last, segs, *a, **k: (sweeps.append(len(segs)), real_last(last, segs, *a, **k))[1])
    grads = {}
    for defer in ("1", "0"):
        monkeypatch.setenv("SPLADE_B200_DEFER_DW", defer)
        w = W.clone().requires_grad_(True)
        b = bias.clone().requires_grad_(True)
        hs = [h.clone().requires_grad_(True) for h, _ in ins]
        sweeps.clear()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            reps = [splade_b200.splade_head(h, w, b, m)[0] for h, (_, m) in zip(hs, ins)]
            loss = sum((r * r).sum() * (i + 1) for i, r in enumerate(reps))
        loss.backward()
        torch.cuda.synchronize()
        assert sweeps == ([3] if defer == "1" else [1, 1, 1])
        grads[defer] = [w.grad, b.grad] + [h.grad for h in hs]
    for a, r in zip(grads["1"], grads["0"]):
        torch.testing.assert_close(a, r, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
# packed (variable-length) token layout: splade_head_fwd_mode
# ---------------------------------------------------------------------------------------------------------
def _both_layouts(hidden, W, bias, mask, want_tokw=True):
    plans, outs = [], []
    for mode in (0, 1):
        plan = torch.full((4,), -7, dtype=torch.int32, device=hidden.device)
        outs.append(run_head(hidden, W, bias, mask, want_tokw, pack_mode=mode, plan_out=plan))
        plans.append(plan.cpu().tolist())
    return outs, plans


@pytest.mark.parametrize("B,S,H,V,kw", [
    (1, 1, 64, 1, {}),
    (3, 7, 64, 130, {}),                       # 21 tokens: one partial packed tile holding three documents
    (8, 64, 128, 50000, {}),                   # config 1 head shape, ragged
    (5, 100, 64, 777, {"holes": True}),        # non-suffix mask: positions come from the column descriptors
    (9, 250, 192, 1000, {}),                   # documents straddle packed tiles (merge keys)
    (2, 300, 64, 260, {}),                     # S > 256
    (2, 700, 128, 200, {"bias_offset": -1.0}),
    (4, 256, 768, 2048, {"bias_offset": -2.0}),
    (40, 33, 64, 300, {}),                     # many short documents per tile, several boundaries per window
])
def test_packed_layout_bit_identical_to_padded(cuda_device, B, S, H, V, kw):
    """The packed tiling computes the same bf16 logits (same K order per column), so sparse_repr, argmax and
    token_weights must be BIT-identical to the padded tiling, and both must match the oracle."""
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=B * 1000 + S + 1, device=cuda_device, **kw)
    if B >= 3:
        mask[1] = 0                            # a document with no real token between two real ones
    for want_tokw in (True, False):
        (pad, pk), plans = _both_layouts(hidden, W, bias, mask, want_tokw)
        assert plans[0][0] == 0 and plans[1][0] == 1, plans
        assert plans[1][1] == int(mask.sum()), plans
        assert torch.equal(pad[0], pk[0]), "sparse_repr differs between the padded and the packed tiling"
        assert torch.equal(pad[1], pk[1]), "argmax differs between the padded and the packed tiling"
        if want_tokw:
            assert torch.equal(pad[2], pk[2]), "token_weights differ between the padded and the packed tiling"
    _check_forward(hidden, W, bias, mask, what=f"packed B{B} S{S} H{H} V{V}", pack_mode=1)


def test_packed_layout_golden(cuda_device, golden):
    for name in ("head_bf16", "head_sparse", "head_s300"):
        g = golden(name)
        t = {k: torch.from_numpy(v).to(cuda_device) for k, v in g.items()}
        rep, arg, tokw = run_head(t["hidden"], t["W"], t["bias"], t["mask"], pack_mode=1)
        near = O.near_tie_mask(g["hidden"], g["W"], g["bias"], g["mask"], mode="bf16")
        assert_argmax_parity(arg.cpu().numpy(), g["argmax"], near, name)
        assert_weights_parity(rep.cpu().numpy(), g["sparse_repr"], name)
        assert_weights_parity(tokw.cpu().numpy(), g["token_weights"], name + " token_weights", max_flip_frac=1e-3)


def test_device_side_plan_picks_the_cheaper_tiling(cuda_device):
    """pack_mode=-1: no host sync; the plan kernel keeps the padded tiling for full batches and switches to the
    packed one for query-like batches (a handful of real tokens in S=64).  Either way the results are the same."""
    B, S, H, V = 64, 64, 128, 3000
    hidden, W, bias, _ = make_head_inputs(B, S, H, V, seed=77, device=cuda_device)
    full = torch.ones(B, S, dtype=torch.int32, device=cuda_device)
    short = query_like_mask(B, S, 4, 12, seed=5, device=cuda_device)
    for mask, want_mode in ((full, 0), (short, 1)):
        plan = torch.zeros(4, dtype=torch.int32, device=cuda_device)
        rep, arg, tokw = run_head(hidden, W, bias, mask, True, pack_mode=-1, plan_out=plan)
        assert plan.cpu().tolist()[0] == want_mode and plan.cpu().tolist()[1] == int(mask.sum())
        rep0, arg0, tokw0 = run_head(hidden, W, bias, mask, True, pack_mode=0)
        assert torch.equal(rep, rep0) and torch.equal(arg, arg0) and torch.equal(tokw, tokw0)
    # the default entry point (splade_head_fwd, policy from the environment: auto) takes the same decision
    rep, arg, tokw = run_head(hidden, W, bias, short, True)
    assert torch.equal(rep, rep0) and torch.equal(arg, arg0) and torch.equal(tokw, tokw0)


def test_packed_layout_full_size_queries(cuda_device):
    """BASELINE config 3 shape (512 queries x S=64, hidden 768 -> 50 000 terms) with realistic query lengths:
    the packed tiling works on ~1/5 of the columns and must reproduce the padded result bit for bit."""
    B, S, H, V = 512, 64, 768, 50000
    hidden, W, bias, _ = make_head_inputs(B, S, H, V, seed=8, device=cuda_device, ragged=False)
    mask = query_like_mask(B, S, 4, 24, seed=9, device=cuda_device)
    (pad, pk), plans = _both_layouts(hidden, W, bias, mask, want_tokw=True)
    assert plans[1][0] == 1 and plans[1][1] == int(mask.sum())
    assert torch.equal(pad[0], pk[0]) and torch.equal(pad[1], pk[1]) and torch.equal(pad[2], pk[2])
    lengths = mask.sum(1)
    assert bool((pk[1] < lengths[:, None]).all())


def test_cta_pair_kernel_passes_the_same_parity_suite(cuda_device):
    """SPLADE_B200_CTA_PAIR=1 selects the cta_group::2 kernel (two CTAs of a cluster share one 256 x N tile).  The
    switch is read once per process, so the forward parity tests are re-run in a child process."""
    import os
    import subprocess
    import sys
    if os.environ.get("SPLADE_B200_CTA_PAIR") == "1":
        pytest.skip("already running in CTA-pair mode")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPLADE_B200_CTA_PAIR="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_head.py", "-m", "gpu", "-q", "-x", "-k",
                        "forward_vs_oracle or golden or headline or all_padding or properties"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_host_buffer_entry_point(cuda_device):
    """splade_head_fwd_host: pinned host hidden/mask in, pinned host sparse_repr/argmax out, one C-ABI call on a
    side stream; same results as the device-resident call."""
    import splade_b200
    B, S, H, V = 6, 40, 128, 3000
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=21, device=cuda_device)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, False)
    slot = splade_b200.HostEncodeSlot(B, S, H, V, cuda_device)
    h_host = h16.cpu().pin_memory()
    m_host = mask.cpu().pin_memory()
    rep_host = torch.empty(B, V, dtype=torch.float32).pin_memory()
    arg_host = torch.empty(B, V, dtype=torch.int32).pin_memory()
    st = torch.cuda.Stream(cuda_device)
    st.wait_stream(torch.cuda.current_stream(cuda_device))
    splade_b200.head_forward_host(h_host, m_host, w16, bias, slot, rep_host, arg_host, st)
    st.synchronize()
    assert torch.equal(rep_host, rep.cpu()) and torch.equal(arg_host, arg.cpu())
