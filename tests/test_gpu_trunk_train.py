"""-m gpu: the native trunk's TRAINING path (NativeTrunkTrain: native forward, hand-driven backward) against
transformers' ModernBertModel with torch.autograd under autocast(bf16) - the trunk the reference trains through
(src/model/splade_modern.py:69-73, src/train/cli/train_v33_ddp.py:337-364) - on a 4-layer model with both layer
types, ragged right-padded batches; row-wise backward kernels against torch restatements."""
import copy

import pytest
import torch
import torch.nn.functional as F

from oracle import torch_reference as TR

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def _tiny(dev, layers=4, hidden=256, heads=4, vocab=3000, seed=0):
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=hidden * 3 // 2,
                           num_hidden_layers=layers, num_attention_heads=heads, pad_token_id=vocab - 1, bos_token_id=0,
                           eos_token_id=1, cls_token_id=0, sep_token_id=1, max_position_embeddings=512,
                           global_attn_every_n_layers=2, local_attention=128)
    torch.manual_seed(seed)
    m = AutoModelForMaskedLM.from_config(cfg).to(dev)
    with torch.no_grad():  # LayerNorm weights away from 1 so that dgamma / gamma scaling are exercised
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    return m


def _batch(B, S, vocab, dev, seed, full=False):
    g = torch.Generator().manual_seed(seed)
    lens = torch.full((B,), S) if full else torch.randint(max(1, S // 8), S + 1, (B,), generator=g)
    lens[0] = S
    ids = torch.randint(6, vocab - 1, (B, S), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, vocab - 1))
    return ids.to(dev), mask.to(dev)


def test_ln_bwd_kernel(cuda_device):
    from splade_b200 import trunk_train as TT
    g = torch.Generator().manual_seed(3)
    rows, H = 517, 768
    x = (torch.randn(rows, H, generator=g) * 2 + 0.3).to(cuda_device).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(H, generator=g)).to(cuda_device).requires_grad_(True)
    dy = torch.randn(rows, H, generator=g).to(cuda_device)
    y = F.layer_norm(x, (H,), gamma, None, 1e-5)
    rx, rg = torch.autograd.grad(y, (x, gamma), dy)
    base = torch.randn(rows, H, generator=g).to(cuda_device)
    dx = base.clone()
    dgam = TT._ln_bwd(dy.contiguous(), x.detach(), gamma.detach(), 1e-5, dx, accumulate=True)   # fp32 dy, accumulate
    torch.testing.assert_close(dx - base, rx, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dgam, rg, rtol=1e-4, atol=1e-3)
    dy16 = dy.to(torch.bfloat16)
    dx2 = torch.empty_like(base)
    dgam2 = TT._ln_bwd(dy16, x.detach(), gamma.detach(), 1e-5, dx2, accumulate=False)             # bf16 dy, overwrite
    rx2, rg2 = torch.autograd.grad(F.layer_norm(x, (H,), gamma, None, 1e-5), (x, gamma), dy16.float())
    torch.testing.assert_close(dx2, rx2, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(dgam2, rg2, rtol=1e-4, atol=1e-3)
    assert torch.equal(dgam2, TT._ln_bwd(dy16, x.detach(), gamma.detach(), 1e-5, dx2, accumulate=False))


def test_geglu_bwd_kernel(cuda_device):
    from splade_b200 import trunk_train as TT
    g = torch.Generator().manual_seed(4)
    rows, I = 300, 1152
    h = torch.randn(rows, 2 * I, generator=g).to(torch.bfloat16).to(cuda_device).requires_grad_(True)
    dgg = torch.randn(rows, I, generator=g).to(torch.bfloat16).to(cuda_device)
    a, gate = h.chunk(2, dim=-1)
    out = F.gelu(a) * gate                      # bf16 ops, as ModernBertMLP under autocast
    (ref,) = torch.autograd.grad(out, h, dgg)
    got = TT._geglu_bwd(h.detach().contiguous(), dgg.contiguous())
    bad = (got.float() - ref.float()).abs() > 2.0 ** -6 * ref.float().abs() + 1e-3
    assert float(bad.float().mean()) <= 1e-3, int(bad.sum())
    assert _rel(got, ref) <= 5e-3


@pytest.mark.parametrize("full", [False, True])
def test_trunk_train_forward_and_gradients_vs_transformers(cuda_device, full):
    import splade_b200
    mlm = _tiny(cuda_device)
    ref = copy.deepcopy(mlm)
    trunk = splade_b200.NativeTrunkTrain(mlm)
    ids, mask = _batch(6, 72, 3000, cuda_device, seed=5, full=full)
    proj = torch.randn(6, 72, 256, generator=torch.Generator().manual_seed(6)).to(cuda_device)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ours = trunk(ids, mask)
        theirs = ref.model(input_ids=ids, attention_mask=mask)[0]
    valid = mask.bool()[..., None]
    assert ours.dtype == torch.bfloat16 and ours.shape == theirs.shape
    assert float(ours.float().masked_select(~valid).abs().max() if (~valid).any() else 0.0) == 0.0
    d = (ours.float() - theirs.float()) * valid
    assert float(d.abs().mean()) <= 1e-2 and float(d.abs().max()) <= 0.25
    # one scalar through both, gradients of EVERY trunk parameter
    (ours.float() * proj * valid).sum().backward()
    (theirs.float() * proj * valid).sum().backward()
    mine = dict(mlm.named_parameters())
    checked = 0
    for name, rp in ref.named_parameters():
        if rp.grad is None:
            continue
        g, rg = mine[name].grad, rp.grad
        assert g is not None, name
        scale = float(rg.abs().max())
        # the reference's own bf16 backward rounds at every Linear; tolerance as for the head (2e-2 L2-relative),
        # a little wider for the element-wise maximum across 4 layers of accumulated bf16 noise
        assert _rel(g, rg) <= 3e-2, (name, _rel(g, rg))
        assert float((g - rg).abs().max()) <= 6e-2 * scale + 1e-7, (name, float((g - rg).abs().max()), scale)
        checked += 1
    assert checked >= 2 + 4 * 5 + 3  # embeddings, 4 layers, final norm


def test_model_trains_through_the_native_trunk(cuda_device):
    """SPLADEModernBERT.forward with use_native_trunk(train=True): loss and every gradient vs the reference module."""
    import splade_b200
    from tests.model_util import tiny_modernbert_dir
    d = tiny_modernbert_dir(vocab_size=5000, hidden=256, heads=4, layers=3, bias_offset=-0.3)
    ours = splade_b200.SPLADEModernBERT(model_name=d).to(cuda_device).train()
    ref = TR.ReferenceSpladeModule(copy.deepcopy(ours.model)).to(cuda_device).train()
    ours.use_native_trunk(False, train=True)
    q = _batch(8, 16, 5000, cuda_device, 7)
    p = _batch(8, 64, 5000, cuda_device, 8)
    n = _batch(8, 64, 5000, cuda_device, 9)
    loss_fn = splade_b200.SPLADELossV33()
    step = 5000
    lam = [loss_fn._lambda_schedule(step, x) for x in (loss_fn.lambda_q, loss_fn.lambda_d, loss_fn.lambda_neg)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        a, _ = ours(*q)
        pp, _ = ours(*p)
        nn_, _ = ours(*n)
        loss, _ = loss_fn(a, pp, nn_, global_step=step)
        r_loss = TR.v33_loss(ref(*q)[0], ref(*p)[0], ref(*n)[0], *lam)
    loss.backward()
    r_loss.backward()
    assert float(loss) == pytest.approx(float(r_loss), rel=2e-2)
    grads = dict(ours.named_parameters())
    for name, rp in ref.named_parameters():
        g, rg = grads[name].grad, rp.grad
        assert g is not None, name
        assert _rel(g, rg) <= 5e-2, (name, _rel(g, rg))


def test_native_trunk_training_steps_through_ddp_and_accumulation(cuda_device):
    """The reference's call pattern (3 encoder calls + loss, (loss / accum).backward(), clip, AdamW) through
    DistributedDataParallel (world size 1, NCCL) with gradient accumulation 2: the native training trunk against the
    reference module, same init - loss trajectory and the parameters after two optimizer steps."""
    import os

    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    import splade_b200
    from splade_b200.training import StepConfig, synthetic_triplet_batch, train_micro_step
    from tests.model_util import tiny_modernbert_dir
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda_device)
    try:
        d = tiny_modernbert_dir(vocab_size=5000, hidden=256, heads=4, layers=3, bias_offset=-0.6)
        ours = splade_b200.SPLADEModernBERT(model_name=d).to(cuda_device)
        ours.use_native_trunk(False, train=True)
        ref = TR.ReferenceSpladeModule(copy.deepcopy(ours.model)).to(cuda_device)
        init = {n: p.detach().clone() for n, p in ours.named_parameters()}
        traj = {}
        for tag, model in (("ours", ours), ("ref", ref)):
            ddp = DDP(model, device_ids=[0], broadcast_buffers=False, find_unused_parameters=False)
            opt = torch.optim.AdamW(ddp.parameters(), lr=5e-5, weight_decay=0.01)
            if tag == "ours":
                loss_fn = splade_b200.SPLADELossV33()
            else:
                sched = splade_b200.SPLADELossV33()

                def loss_fn(anchor_repr, positive_repr, negative_repr, global_step, **kw):
                    lam = [sched._lambda_schedule(global_step, x)
                           for x in (sched.lambda_q, sched.lambda_d, sched.lambda_neg)]
                    return TR.v33_loss(anchor_repr, positive_repr, negative_repr, *lam), {}
            gs, losses = 0, []
            cfg = StepConfig(gradient_accumulation_steps=2)
            for i in range(4):
                batch = synthetic_triplet_batch(4, 16, 48, 1, 5000, 4999, seed=100 + i)
                loss, _, gs = train_micro_step(ddp, loss_fn, opt, None, batch, i, gs, cfg, cuda_device)
                losses.append(float(loss))
            traj[tag] = losses
            del ddp, opt
        assert traj["ours"] == pytest.approx(traj["ref"], rel=2e-2), traj
        # AdamW's first steps move every element by ~lr in the direction of its gradient's sign: the two runs' UPDATES
        # must point the same way (cosine 1 - 2 x the fraction of sign flips; flips only where the gradient is
        # smaller than the reference's own bf16 noise)
        for (name, p), (_, rp) in zip(ours.named_parameters(), ref.named_parameters()):
            du, dr = (p.detach() - init[name]).flatten().float(), (rp.detach() - init[name]).flatten().float()
            if float(dr.norm()) == 0.0:
                assert float(du.norm()) == 0.0, name
                continue
            cos = float(torch.dot(du, dr) / (du.norm() * dr.norm()).clamp_min(1e-30))
            assert cos >= 0.7, (name, cos)
    finally:
        dist.destroy_process_group()


def test_fused_encoder_calls_give_the_same_vectors_and_gradients(cuda_device):
    """StepConfig.fuse_encoder_calls: one encoder call for queries + positives + negatives (padded to a common
    length) against three calls - same sparse vectors (bit for bit through the packed native trunk: a token's row
    does not depend on its neighbours) and the same parameter gradients up to summation order."""
    import splade_b200
    from splade_b200.training import concat_padded
    from tests.model_util import tiny_modernbert_dir
    d = tiny_modernbert_dir(vocab_size=5000, hidden=256, heads=4, layers=3, bias_offset=-0.3)
    model = splade_b200.SPLADEModernBERT(model_name=d).to(cuda_device).train()
    model.use_native_trunk(False, train=True)
    q, p = _batch(8, 16, 5000, cuda_device, 17), _batch(8, 64, 5000, cuda_device, 18)
    loss_fn = splade_b200.SPLADELossV33()

    def run(fused):
        model.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if fused:
                ids, mask = concat_padded((q[0], p[0]), (q[1], p[1]))
                a, pp = model(ids, mask)[0].split([8, 8])
            else:
                a, pp = model(*q)[0], model(*p)[0]
            loss, _ = loss_fn(a, pp, pp.flip(0), global_step=100)
        loss.backward()
        return a.detach(), pp.detach(), {n: t.grad.clone() for n, t in model.named_parameters() if t.grad is not None}

    a1, p1, g1 = run(False)
    a2, p2, g2 = run(True)
    assert torch.equal(a1, a2) and torch.equal(p1, p2)
    for n in g1:
        assert _rel(g2[n], g1[n]) <= 1e-3, (n, _rel(g2[n], g1[n]))
