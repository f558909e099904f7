"""-m gpu: native variable-length ModernBERT trunk (splade_trunk_encode and its building blocks) against plain
torch restatements of the same arithmetic and against transformers' ModernBertForMaskedLM on the same device."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.to(torch.bfloat16)


def _close(got, ref, what, atol, rtol=2e-2, frac=1e-3):
    got, ref = got.float(), ref.float()
    bad = (got - ref).abs() > atol + rtol * ref.abs()
    assert float(bad.float().mean()) <= frac, (what, int(bad.sum()), bad.numel(), float((got - ref).abs().max()))


@pytest.mark.parametrize("rows,K,N", [(300, 256, 512), (128, 64, 256), (1000, 768, 768), (257, 1152, 768)])
def test_trunk_gemm_epilogues(cuda_device, rows, K, N):
    import splade_b200
    from splade_b200 import _lib as L
    g = torch.Generator().manual_seed(rows + K)
    A = _bf(torch.randn(rows, K, generator=g)).to(cuda_device)
    W = _bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(cuda_device)
    acc = A.float() @ W.float().t()
    ref = _bf(acc)
    out = splade_b200.trunk_gemm(L.EPI_BF16, A, W)
    _close(out, ref, "bf16", atol=2e-2)
    out = splade_b200.trunk_gemm(L.EPI_GELU, A, W)
    _close(out, _bf(torch.nn.functional.gelu(ref.float())), "gelu", atol=2e-2)
    x0 = torch.randn(rows, N, generator=g).to(cuda_device)
    x = x0.clone()
    splade_b200.trunk_gemm(L.EPI_RESID, A, W, out=x)
    _close(x, x0 + ref.float(), "resid", atol=2e-2)
    # GeGLU: W = [input rows | gate rows]
    Wg = _bf(torch.randn(2 * N, K, generator=g) / math.sqrt(K)).to(cuda_device)
    a2 = _bf(A.float() @ Wg.float().t())
    refg = _bf(_bf(torch.nn.functional.gelu(a2[:, :N].float())).float() * a2[:, N:].float())
    out = splade_b200.trunk_gemm(L.EPI_GEGLU, A, Wg)
    _close(out, refg, "geglu", atol=2e-2)


def _pack_rows(x_pad, lens, rows):
    """[B,S,C] padded -> [rows,C] in the packed layout: documents start at multiples of 8 rows, fillers behind a
    document replicate its last token."""
    out = torch.zeros(rows, x_pad.shape[-1], dtype=x_pad.dtype, device=x_pad.device)
    starts, o = [], 0
    for b, L in enumerate(lens):
        starts.append(o)
        if L:
            L8 = (L + 7) // 8 * 8
            out[o:o + L] = x_pad[b, :L]
            out[o + L:o + L8] = x_pad[b, L - 1]
            o += L8
    return out, starts


def _rope(x, theta):
    """apply_rotary_pos_emb on [L, nh, 64] (positions 0..L-1), fp32 math on bf16 inputs, rounded back to bf16."""
    L = x.shape[0]
    inv = 1.0 / (theta ** (torch.arange(0, 64, 2, device=x.device).float() / 64))
    ang = torch.arange(L, device=x.device).float()[:, None] * inv[None, :]
    cos = torch.cat([ang.cos(), ang.cos()], -1)[:, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], -1)[:, None, :]
    rot = torch.cat([-x[..., 32:], x[..., :32]], -1)
    return _bf(x * cos + rot * sin).float()


@pytest.mark.parametrize("window,theta,S,lens", [
    (-1, 0.0, 200, [200, 1, 0, 77, 130, 64]),            # every tile fits 256 keys: tcgen05 kernel
    (64, 0.0, 200, [200, 1, 0, 77, 130, 64]),
    (-1, 160000.0, 200, [200, 1, 0, 77, 130, 64]),
    (64, 10000.0, 200, [200, 1, 0, 77, 130, 64]),
    (-1, 0.0, 600, [600, 256, 257, 3, 0, 411]),          # global over > 256 tokens: mma.sync kernel for those docs
    (64, 10000.0, 600, [600, 256, 257, 3, 0, 411]),      # sliding window: tcgen05 kernel for any length
    (-1, 0.0, 48, [48, 1, 17, 40, 8, 33]),               # query-length batch
])
def test_trunk_attention_vs_torch(cuda_device, window, theta, S, lens):
    import splade_b200
    B, nh = len(lens), 4
    H = nh * 64
    g = torch.Generator().manual_seed(11 + window + S)
    mask = (torch.arange(S)[None, :] < torch.tensor(lens)[:, None]).to(torch.int32).to(cuda_device)
    qkv = _bf(torch.randn(B, S, 3 * H, generator=g)).to(cuda_device)
    rows = splade_b200.load_library().splade_packed_rows(B, S)
    packed, starts = _pack_rows(qkv, lens, rows)
    out, meta = splade_b200.trunk_attention(packed.contiguous(), mask, nh, window, rope_theta=theta)
    for b, L in enumerate(lens):
        if L == 0:
            continue
        x = qkv[b, :L].float().view(L, 3, nh, 64)
        if theta > 0:
            x = torch.stack([_rope(x[:, 0], theta), _rope(x[:, 1], theta), x[:, 2]], 1)
        q, k, v = (x[:, i].transpose(0, 1) for i in range(3))           # [nh, L, 64]
        s = q @ k.transpose(1, 2) * 0.125
        if window >= 0:
            i = torch.arange(L, device=cuda_device)
            s = s.masked_fill(((i[:, None] - i[None, :]).abs() > window)[None], float("-inf"))
        ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(L, H)
        got = out[starts[b]:starts[b] + L]
        _close(got, ref, f"attention doc {b} window {window}", atol=2e-2)
        L8 = (L + 7) // 8 * 8
        if L8 > L:  # replicas reproduce the last token's output exactly
            assert torch.equal(out[starts[b] + L:starts[b] + L8], out[starts[b] + L - 1].expand(L8 - L, -1))
        m = meta[starts[b]:starts[b] + L].cpu().numpy()
        assert np.array_equal(m & 0xFFFF, np.arange(L)) and np.all(((m >> 16) & 0x7FFF) == b)


def _tiny_mlm(device, layers=4):
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=3000, hidden_size=256, intermediate_size=384, num_hidden_layers=layers,
                           num_attention_heads=4, pad_token_id=2999, bos_token_id=0, eos_token_id=1, cls_token_id=0,
                           sep_token_id=1, max_position_embeddings=1024, global_attn_every_n_layers=3,
                           local_attention=128)
    torch.manual_seed(0)
    mlm = AutoModelForMaskedLM.from_config(cfg).to(device).eval()
    with torch.no_grad():
        mlm.decoder.bias.normal_(0.0, 0.02)
        for n, p_ in mlm.named_parameters():
            if n.endswith("norm.weight"):
                p_.add_(0.1 * torch.randn_like(p_))   # LayerNorm weights away from 1 so that they are exercised
    return mlm, cfg


def test_native_trunk_vs_transformers(cuda_device):
    """Whole network on a small ModernBERT (4 layers: global, local, local, global): the packed head input and the
    pooled sparse vectors against transformers' own forward under autocast(bf16) on the same device."""
    import splade_b200
    mlm, cfg = _tiny_mlm(cuda_device)
    B, S = 7, 160
    g = torch.Generator().manual_seed(3)
    lens = torch.tensor([160, 5, 33, 97, 128, 1, 150])
    ids = torch.randint(6, cfg.vocab_size - 1, (B, S), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64)
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, cfg.pad_token_id)).to(cuda_device)
    mask = mask.to(cuda_device)
    trunk = splade_b200.NativeTrunk(mlm)
    rep, dbg = trunk.encode(ids, mask, debug=True)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref_hidden = mlm.head(mlm.model(input_ids=ids, attention_mask=mask)[0]).float()
        ref_rep, _ = splade_b200.splade_head(ref_hidden, mlm.decoder.weight, mlm.decoder.bias, mask,
                                             want_token_weights=False)
    plan = dbg["plan"].cpu().tolist()
    assert plan[0] == 1 and plan[4] == int(lens.sum())
    meta = dbg["meta"].cpu().numpy().astype(np.int64)
    T = plan[1]
    doc, pos = (meta[:T] >> 16) & 0x7FFF, meta[:T] & 0xFFFF
    got = dbg["hidden_packed"][:T].float()
    want = ref_hidden[torch.from_numpy(doc).to(cuda_device), torch.from_numpy(pos).to(cuda_device)]
    err = (got - want).abs()
    # bf16 Linear outputs through 4 layers: a few 1e-3 of noise on unit-scale (post-LayerNorm) activations
    assert float(err.mean()) < 1e-2 and float(err.max()) < 0.25, (float(err.mean()), float(err.max()))
    cos = torch.nn.functional.cosine_similarity(rep, ref_rep, dim=1)
    assert float(cos.min()) > 0.999, cos
    np.testing.assert_allclose(rep.cpu().numpy(), ref_rep.cpu().numpy(), atol=3e-2, rtol=3e-2)
    assert float(rep[torch.where(lens == 0)[0]].abs().sum()) == 0.0


def _hf_vs_native(cuda_device, lens, S, hole_at=None, seed=3):
    import splade_b200
    mlm, cfg = _tiny_mlm(cuda_device)
    B = len(lens)
    g = torch.Generator().manual_seed(seed)
    lens = torch.tensor(lens)
    ids = torch.randint(6, cfg.vocab_size - 1, (B, S), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64)
    if hole_at is not None:
        mask[:, hole_at] = 0                      # non-suffix masks: positions keep counting across the holes
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, cfg.pad_token_id)).to(cuda_device)
    mask = mask.to(cuda_device)
    rep, dbg = splade_b200.NativeTrunk(mlm).encode(ids, mask, debug=True)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref_hidden = mlm.head(mlm.model(input_ids=ids, attention_mask=mask)[0]).float()
        ref_rep, _ = splade_b200.splade_head(ref_hidden, mlm.decoder.weight, mlm.decoder.bias, mask,
                                             want_token_weights=False)
    plan = dbg["plan"].cpu().tolist()
    assert plan[4] == int(mask.sum())
    meta = dbg["meta"].cpu().numpy().astype(np.int64)
    T = plan[1]
    doc, pos = (meta[:T] >> 16) & 0x7FFF, meta[:T] & 0xFFFF
    assert bool(mask[torch.from_numpy(doc).to(cuda_device), torch.from_numpy(pos).to(cuda_device)].all())
    got = dbg["hidden_packed"][:T].float()
    want = ref_hidden[torch.from_numpy(doc).to(cuda_device), torch.from_numpy(pos).to(cuda_device)]
    err = (got - want).abs()
    assert float(err.mean()) < 1e-2 and float(err.max()) < 0.3, (float(err.mean()), float(err.max()))
    live = mask.sum(1) > 0
    cos = torch.nn.functional.cosine_similarity(rep[live], ref_rep[live], dim=1)
    assert float(cos.min()) > 0.999, cos
    assert float(rep[~live].abs().sum()) == 0.0


def test_native_trunk_long_documents_use_both_attention_kernels(cuda_device):
    """Documents longer than 256 tokens: global layers fall to the mma.sync kernel for those documents while the
    short ones of the same batch (and every sliding-window layer) run on the tcgen05 kernel."""
    _hf_vs_native(cuda_device, [700, 300, 257, 9, 256, 0], 700)


def test_native_trunk_with_holes_in_the_mask(cuda_device):
    _hf_vs_native(cuda_device, [90, 12, 64, 33], 96, hole_at=5, seed=6)


def test_native_trunk_is_padding_and_batch_invariant(cuda_device):
    """A document's vector does not depend on its neighbours, on the padded width, or on what sits at padded ids."""
    import splade_b200
    mlm, cfg = _tiny_mlm(cuda_device, layers=3)
    trunk = splade_b200.NativeTrunk(mlm)
    g = torch.Generator().manual_seed(4)
    B, S = 5, 96
    lens = torch.tensor([96, 17, 40, 8, 64])
    ids = torch.randint(6, cfg.vocab_size - 1, (B, S), generator=g).to(cuda_device)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64).to(cuda_device)
    rep = trunk.encode(ids, mask)
    rep2 = trunk.encode(ids[[2, 0]].contiguous(), mask[[2, 0]].contiguous())
    assert torch.equal(rep2[0], rep[2]) and torch.equal(rep2[1], rep[0])
    wide_ids = torch.cat([ids, torch.randint(6, 100, (B, 32), generator=g).to(cuda_device)], 1)
    wide_mask = torch.cat([mask, torch.zeros(B, 32, dtype=mask.dtype, device=cuda_device)], 1)
    assert torch.equal(trunk.encode(wide_ids, wide_mask), rep)


def test_dropin_encoder_and_bulk_encode_use_the_native_trunk(cuda_device, tmp_path):
    """``SPLADEModernBERT.use_native_trunk`` / ``bulk_encode(native_trunk=True)``: same interface, vectors that agree
    with the transformers-trunk path to bf16 noise, identical ``state_dict`` keys, and ``forward`` (training) keeps
    the reference trunk."""
    import splade_b200
    mlm, cfg = _tiny_mlm(cuda_device, layers=3)
    mlm.save_pretrained(tmp_path)
    model = splade_b200.SPLADEModernBERT(model_name=str(tmp_path)).to(cuda_device).eval()
    keys = set(model.state_dict().keys())
    g = torch.Generator().manual_seed(8)
    N, S = 24, 64
    lens = torch.randint(3, S + 1, (N,), generator=g)
    ids = torch.randint(6, cfg.vocab_size - 1, (N, S), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ref = model.encode(ids.to(cuda_device), mask.to(cuda_device))
        model.use_native_trunk(True)
        nat = model.encode(ids.to(cuda_device), mask.to(cuda_device))
    assert set(model.state_dict().keys()) == keys
    assert float(torch.nn.functional.cosine_similarity(nat, ref, dim=1).min()) > 0.999
    # training path: autograd through the reference trunk even when the native trunk is enabled for encode
    rep, tokw = model(ids[:4].to(cuda_device), mask[:4].to(cuda_device))
    rep.sum().backward()
    assert model.model.decoder.weight.grad is not None and tokw.shape == (4, S)
    model.use_native_trunk(False)
    i0, v0, c0, _ = splade_b200.bulk_encode(model, ids, mask, 16, batch_size=8, native_trunk=False)
    i1, v1, c1, _ = splade_b200.bulk_encode(model, ids, mask, 16, batch_size=8, native_trunk=True)
    assert model._native_trunk_enabled is False
    assert torch.equal(c0, c1)
    overlap = np.mean([len(set(a.tolist()) & set(b.tolist())) / 16.0 for a, b in zip(i0.cpu(), i1.cpu())])
    assert overlap > 0.9, overlap
    np.testing.assert_allclose(v1[:, 0].cpu().numpy(), v0[:, 0].cpu().numpy(), rtol=3e-2, atol=3e-2)


def test_native_trunk_real_shape_vs_transformers_and_fp32(cuda_device):
    """The native trunk at the REAL A.X-Encoder-base shape (22 layers, hidden 768, 12 heads, intermediate 1152,
    vocab 50 000; the shape every native-trunk number in bench.py / profiles rests on), B = 64 x S = 256 where the
    +-64 sliding window actually bites, full-length and ragged batches.  Two oracles, thresholds stated here:

      (1) transformers' own forward under autocast(bf16) + the fused head (the drop-in default path):
          per-document cosine >= 0.9999, max |dw| <= 2e-2, mean top-64 overlap >= 0.98, worst document >= 0.90
      (2) the reference's fp32 inference chain (benchmark/encoders.py:309-318, no autocast): the native trunk must
          be no further from it than the reference's OWN autocast path is (cosine and top-64 overlap within 1e-4 /
          0.01 of that path's) - bf16 Linear outputs through 22 layers are the reference's rounding noise too.
    """
    import splade_b200
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    V, H = 50000, 768
    cfg = ModernBertConfig(vocab_size=V, hidden_size=H, intermediate_size=1152, num_hidden_layers=22,
                           num_attention_heads=12, pad_token_id=V - 1, bos_token_id=0, eos_token_id=1, cls_token_id=0,
                           sep_token_id=1, max_position_embeddings=16384, global_attn_every_n_layers=3,
                           local_attention=128)
    torch.manual_seed(0)
    mlm = AutoModelForMaskedLM.from_config(cfg).to(cuda_device).eval()
    with torch.no_grad():
        mlm.decoder.bias.normal_(0.0, 0.02)
    trunk = splade_b200.NativeTrunk(mlm)
    B, S = 64, 256
    g = torch.Generator().manual_seed(5)

    def topk_overlap(a, b, k=64):
        ta, tb = a.topk(k, dim=1).indices.tolist(), b.topk(k, dim=1).indices.tolist()
        return [len(set(x) & set(y)) / k for x, y in zip(ta, tb)]

    for ragged in (False, True):
        lens = torch.randint(32, S + 1, (B,), generator=g) if ragged else torch.full((B,), S)
        if ragged:
            lens[0], lens[1] = S, 1
        ids = torch.randint(6, V - 1, (B, S), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64)
        ids = torch.where(mask.bool(), ids, torch.full_like(ids, cfg.pad_token_id)).to(cuda_device)
        mask = mask.to(cuda_device)
        native = trunk.encode(ids, mask)
        with torch.no_grad():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                hid = mlm.head(mlm.model(input_ids=ids, attention_mask=mask)[0]).float()
                auto, _ = splade_b200.splade_head(hid, mlm.decoder.weight, mlm.decoder.bias, mask,
                                                  want_token_weights=False)
            logits = mlm(input_ids=ids, attention_mask=mask).logits            # fp32, no autocast
            fp32 = (torch.log1p(torch.relu(logits)) * mask.unsqueeze(-1).float()).max(dim=1).values
            del logits
        what = "ragged" if ragged else "full"
        # (1) against the drop-in default (transformers trunk under autocast + fused head)
        cos = torch.nn.functional.cosine_similarity(native, auto, dim=1)
        ov = topk_overlap(native, auto)
        assert float(cos.min()) >= 0.9999, (what, float(cos.min()))
        assert float((native - auto).abs().max()) <= 2e-2, (what, float((native - auto).abs().max()))
        assert sum(ov) / len(ov) >= 0.98 and min(ov) >= 0.90, (what, sum(ov) / len(ov), min(ov))
        # (2) against fp32: no worse than the reference's own autocast path
        cos_n = float(torch.nn.functional.cosine_similarity(native, fp32, dim=1).min())
        cos_a = float(torch.nn.functional.cosine_similarity(auto, fp32, dim=1).min())
        ov_n, ov_a = topk_overlap(native, fp32), topk_overlap(auto, fp32)
        assert cos_n >= cos_a - 1e-4, (what, cos_n, cos_a)
        assert sum(ov_n) / len(ov_n) >= sum(ov_a) / len(ov_a) - 0.01, (what, sum(ov_n) / B, sum(ov_a) / B)
        np.testing.assert_allclose(native.cpu().numpy(), fp32.cpu().numpy(), atol=2e-2, rtol=2e-2)
