"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE in the build container.

Run (here, where /root/reference exists; it does not exist on the GPU box):
    python oracle/make_golden.py

The reference modules are loaded by file path (``import src.model`` would pull in sentence_transformers, which
is not installed: SURVEY.md §8(c) gotcha 1).  Nothing is copied from the reference; only its OUTPUTS on seeded
inputs are stored.  Cases:

  g1_pool        SURVEY §8(c) G1: reference forward/backward on hand-written logits (stub MLM model)
  g2_loss        SURVEY §8(c) G2: SPLADELossV33 defaults at steps 0 / 10000 / 20000, with gradients
  g3_multineg    SURVEY §8(c) G3: k=2 negatives + MarginMSE
  head_bf16      reference forward/backward with a stub MLM model whose logits are the autocast-style
                 bf16-rounded projection of seeded (hidden, W, bias); ragged right-padded mask
  head_s300      same with S=300 (documents longer than one 256-token tile) and a non-suffix mask
  c1_tiny        BASELINE config 1: real tiny random-init ModernBERT (2 layers, hidden 128), fp32 CPU, with the
                 head input captured by a forward hook (vocab 1111 to keep the fixture small)
  loss_rand_*    SPLADELossV33 on random sparse-ish reps: k=1, k=3, KL-KD, MarginMSE, temperature != 1
  emit           NeuralSparseEncoderV33._encode_batch (benchmark/encoders.py:309-345) with a stub model and a
                 synthetic token table: top_k = None / 5 / 64, ties and block-listed ids included
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _load(name: str, path: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    sm = _load("ref_splade_modern", f"{REF}/src/model/splade_modern.py")
    ls = _load("ref_losses", f"{REF}/src/model/losses.py")
    return sm.SPLADEModernBERT, ls.SPLADELossV33


def bf16r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


class _StubMLM(torch.nn.Module):
    """Stands in for HF ModernBertForMaskedLM inside the reference class: returns given logits."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn
        self.config = types.SimpleNamespace(vocab_size=0, hidden_size=0)

    def forward(self, input_ids=None, attention_mask=None):
        return types.SimpleNamespace(logits=self.fn())


def make_ref_model(SPLADEModernBERT, logits_fn):
    m = SPLADEModernBERT.__new__(SPLADEModernBERT)
    torch.nn.Module.__init__(m)
    m.model_name = "stub"
    m.model = _StubMLM(logits_fn)
    m.config = m.model.config
    m.relu = torch.nn.ReLU()
    return m


def case_g1(SPLADEModernBERT):
    logits = torch.tensor([[[0.5, -1, 0.1], [2, 0, 0.2], [2, -3, 0.3], [1, -0.5, 9]]], requires_grad=True)
    mask = torch.tensor([[1, 1, 1, 0]])
    m = make_ref_model(SPLADEModernBERT, lambda: logits)
    rep, tokw = m(torch.zeros(1, 4, dtype=torch.long), mask)
    rep.sum().backward()
    return dict(logits=logits.detach().numpy(), mask=mask.numpy(), sparse_repr=rep.detach().numpy(),
                token_weights=tokw.detach().numpy(), dlogits=logits.grad.numpy())


def head_case(SPLADEModernBERT, B, S, H, V, seed, lengths=None, holes=False, bias_offset=0.0):
    g = torch.Generator().manual_seed(seed)
    hidden = bf16r(torch.randn(B, S, H, generator=g)).requires_grad_(True)
    W = bf16r(torch.randn(V, H, generator=g) * 0.0176 * (768 / H) ** 0.5).requires_grad_(True)
    bias = bf16r(torch.randn(V, generator=g) * 0.02 + bias_offset).requires_grad_(True)
    if lengths is None:
        lengths = torch.randint(max(1, S // 8), S + 1, (B,), generator=g)
        lengths[0] = S
    mask = (torch.arange(S)[None, :] < lengths[:, None]).long()
    if holes:  # non-suffix mask: knock out a few interior positions
        mask[:, 3] = 0
        mask[0, S // 2] = 0
    grad_out = torch.randn(B, V, generator=g)

    class _RoundBf16(torch.autograd.Function):  # autocast's cast of the logits: straight-through in autograd
        @staticmethod
        def forward(ctx, x):
            return bf16r(x)

        @staticmethod
        def backward(ctx, gout):
            return gout

    def logits_fn():
        return _RoundBf16.apply(torch.nn.functional.linear(hidden, W, bias))

    m = make_ref_model(SPLADEModernBERT, logits_fn)
    rep, tokw = m(torch.zeros(B, S, dtype=torch.long), mask)
    logits = logits_fn().detach()
    scores = torch.log1p(torch.relu(logits)) * mask.unsqueeze(-1).float()
    argmax = scores.max(dim=1).indices
    (rep * grad_out).sum().backward()
    return dict(hidden=hidden.detach().numpy(), W=W.detach().numpy(), bias=bias.detach().numpy(),
                mask=mask.numpy().astype(np.int32), grad_out=grad_out.numpy(),
                sparse_repr=rep.detach().numpy(), token_weights=tokw.detach().numpy(),
                argmax=argmax.numpy().astype(np.int32), d_hidden=hidden.grad.numpy(), dW=W.grad.numpy(),
                dbias=bias.grad.numpy())


def case_c1(SPLADEModernBERT):
    import tempfile

    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=1111, hidden_size=128, intermediate_size=192, num_hidden_layers=2,
                           num_attention_heads=4, pad_token_id=1110, bos_token_id=0, eos_token_id=1,
                           cls_token_id=0, sep_token_id=1, max_position_embeddings=512)
    torch.manual_seed(0)
    hf = AutoModelForMaskedLM.from_config(cfg)
    with torch.no_grad():
        hf.decoder.bias.copy_(torch.randn(1111, generator=torch.Generator().manual_seed(2)) * 0.02)
    with tempfile.TemporaryDirectory() as d:
        hf.save_pretrained(d)
        model = SPLADEModernBERT(model_name=d).eval()
    g = torch.Generator().manual_seed(1)
    B, S = 8, 64
    ids = torch.randint(6, 1110, (B, S), generator=g)
    lengths = torch.randint(S // 8, S + 1, (B,), generator=g)
    lengths[0] = S
    mask = (torch.arange(S)[None, :] < lengths[:, None]).long()
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, 1110))
    captured = {}
    hook = model.model.head.register_forward_hook(lambda mod, inp, out: captured.__setitem__("h", out.detach()))
    with torch.no_grad():
        rep, tokw = model(ids, mask)
    hook.remove()
    return dict(input_ids=ids.numpy(), mask=mask.numpy().astype(np.int32), head_hidden=captured["h"].numpy(),
                W=model.model.decoder.weight.detach().numpy(), bias=model.model.decoder.bias.detach().numpy(),
                sparse_repr=rep.numpy(), token_weights=tokw.numpy())


def sparse_reps(g, n, V, density=0.2, scale=1.0):
    x = torch.rand(n, V, generator=g) * scale
    keep = torch.rand(n, V, generator=g) < density
    return x * keep


def loss_case(SPLADELossV33, *, Bq, V, k, seed, step, ctor=None, kd=False, mmse=False):
    ctor = ctor or {}
    g = torch.Generator().manual_seed(seed)
    q = sparse_reps(g, Bq, V).requires_grad_(True)
    p = sparse_reps(g, Bq, V, density=0.3).requires_grad_(True)
    n = sparse_reps(g, Bq * k, V, density=0.3)
    n = (n.view(Bq, k, V) if k > 1 else n).requires_grad_(True)
    kwargs = {}
    if kd:
        kwargs["teacher_scores"] = torch.randn(Bq, Bq, generator=g) * 2
    if mmse:
        kwargs["teacher_pos_scores"] = torch.rand(Bq, generator=g) * 5
        kwargs["teacher_neg_scores"] = torch.rand(Bq, k, generator=g) * 5 if k > 1 else torch.rand(Bq, generator=g) * 5
    fn = SPLADELossV33(**ctor)
    loss, ld = fn(q, p, n, global_step=step, **kwargs)
    loss.backward()
    out = dict(q=q.detach().numpy(), p=p.detach().numpy(), n=n.detach().numpy(), step=np.int64(step),
               loss=np.float64(loss.item()), dq=q.grad.numpy(), dp=p.grad.numpy(), dn=n.grad.numpy(),
               avg_nonzero=np.array(fn.get_avg_nonzero(), dtype=np.float64))
    for key, val in ld.items():
        out["ld_" + key] = np.float64(val)
    for key, val in ctor.items():
        out["ctor_" + key] = np.float64(val)
    for key, val in kwargs.items():
        out[key] = val.numpy()
    return out


def case_g2(SPLADELossV33):
    out = {}
    qv = [[1, 0, 2, 0], [0, 1, 0, 1]]
    pv = [[1, 1, 0, 0], [0, 2, 0, 1]]
    nv = [[0, 0, 1, 0], [1, 0, 0, 0]]
    for step in (0, 10000, 20000):
        q = torch.tensor(qv, dtype=torch.float32, requires_grad=True)
        p = torch.tensor(pv, dtype=torch.float32, requires_grad=True)
        n = torch.tensor(nv, dtype=torch.float32, requires_grad=True)
        loss, ld = SPLADELossV33()(q, p, n, global_step=step)
        loss.backward()
        out[f"loss_{step}"] = np.float64(loss.item())
        out[f"dq_{step}"], out[f"dp_{step}"], out[f"dn_{step}"] = q.grad.numpy(), p.grad.numpy(), n.grad.numpy()
        for key, val in ld.items():
            out[f"ld_{key}_{step}"] = np.float64(val)
    out.update(q=np.array(qv, np.float32), p=np.array(pv, np.float32), n=np.array(nv, np.float32))
    return out


def case_g3(SPLADELossV33):
    q = torch.tensor([[1, 0, 2, 0], [0, 1, 0, 1]], dtype=torch.float32, requires_grad=True)
    p = torch.tensor([[1, 1, 0, 0], [0, 2, 0, 1]], dtype=torch.float32, requires_grad=True)
    n = torch.tensor([[[0, 0, 1, 0], [0, 1, 0, 0]], [[1, 0, 0, 0], [0, 0, 0, 2]]], dtype=torch.float32,
                     requires_grad=True)
    tp = torch.tensor([0.9, 0.8])
    tn = torch.tensor([[0.2, 0.1], [0.3, 0.4]])
    fn = SPLADELossV33(lambda_margin_mse=0.5, lambda_neg=0.01, flops_warmup_steps=0)
    loss, ld = fn(q, p, n, global_step=5, teacher_pos_scores=tp, teacher_neg_scores=tn)
    loss.backward()
    out = dict(q=q.detach().numpy(), p=p.detach().numpy(), n=n.detach().numpy(), tp=tp.numpy(), tn=tn.numpy(),
               loss=np.float64(loss.item()), dq=q.grad.numpy(), dp=p.grad.numpy(), dn=n.grad.numpy(),
               avg_nonzero=np.array(fn.get_avg_nonzero(), dtype=np.float64))
    for key, val in ld.items():
        out["ld_" + key] = np.float64(val)
    return out


def case_emit():
    """Run the reference's _encode_batch on a stub model + synthetic vocabulary."""
    st = types.ModuleType("sentence_transformers")
    st.SentenceTransformer = object
    sys.modules.setdefault("sentence_transformers", st)
    sys.path.insert(0, REF)
    # benchmark.encoders imports src.model.splade_modern via the package path; pre-seed light stand-ins so that
    # `import src` does not run src/__init__.py (-> teachers -> sentence_transformers models).
    for name in ("src", "src.model", "benchmark"):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(REF, *name.split("."))]
        sys.modules.setdefault(name, pkg)
    enc_mod = _load("ref_bench_encoders", f"{REF}/benchmark/encoders.py")
    Enc = enc_mod.NeuralSparseEncoderV33

    V = 300
    g = torch.Generator().manual_seed(7)
    tokens = [f"tok{i}" for i in range(V - 1)]  # tokenizer.vocab_size = V-1, like 49 999 vs 50 000
    for i in (0, 1, 2, 3, 4, 5, 17, 40):
        tokens[i] = f"[unused{i}]" if i % 2 == 0 else f"<sp{i}>"
    special = {0, 1, 2, 3, 5, V - 1}
    rep = sparse_reps(g, 6, V, density=0.5)
    rep = (rep * 8).round() / 8          # few distinct values -> many ties, like bf16-quantised weights
    rep[1] = 0                           # an empty row
    rep[2, 10:200] = 0.5                 # a long tie run across the k-th position
    rep[3] = (torch.rand(V, generator=g) < 0.01).float() * 0.75  # fewer survivors than k

    class _M(torch.nn.Module):
        def forward(self, input_ids=None, attention_mask=None):
            return rep, None

    enc = Enc.__new__(Enc)
    enc.model = _M()
    enc.special_token_ids = special
    enc._token_lookup = tokens
    out = dict(rep=rep.numpy(), blocklist=np.zeros(V, np.uint8))
    for i in range(V):
        if i in special or i >= len(tokens) or not tokens[i] or tokens[i].startswith(("[", "<")):
            out["blocklist"][i] = 1
    for k in (None, 5, 64):
        dicts = enc._encode_batch({"input_ids": None, "attention_mask": None}, k)
        tag = "all" if k is None else str(k)
        for r, d in enumerate(dicts):
            ids = np.array([tokens.index(t) for t in d.keys()], dtype=np.int64)
            out[f"ids_{tag}_{r}"] = ids
            out[f"w_{tag}_{r}"] = np.array(list(d.values()), dtype=np.float32)
    out["tokens"] = np.array(tokens)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    SPLADEModernBERT, SPLADELossV33 = load_reference()
    cases = {
        "g1_pool": case_g1(SPLADEModernBERT),
        "g2_loss": case_g2(SPLADELossV33),
        "g3_multineg": case_g3(SPLADELossV33),
        "head_bf16": head_case(SPLADEModernBERT, B=5, S=37, H=64, V=300, seed=11),
        "head_sparse": head_case(SPLADEModernBERT, B=4, S=64, H=128, V=391, seed=12, bias_offset=-1.0),
        "head_s300": head_case(SPLADEModernBERT, B=2, S=300, H=64, V=200, seed=13, holes=True),
        "c1_tiny": case_c1(SPLADEModernBERT),
        "loss_rand_k1": loss_case(SPLADELossV33, Bq=6, V=257, k=1, seed=21, step=1234),
        "loss_rand_k3": loss_case(SPLADELossV33, Bq=5, V=130, k=3, seed=22, step=30000,
                                  ctor=dict(lambda_neg=0.02, temperature=2.0)),
        "loss_rand_kd": loss_case(SPLADELossV33, Bq=7, V=96, k=1, seed=23, step=500, kd=True,
                                  ctor=dict(lambda_kd=0.7, kd_temperature=1.5)),
        "loss_rand_mmse": loss_case(SPLADELossV33, Bq=4, V=64, k=2, seed=24, step=0, mmse=True,
                                    ctor=dict(lambda_margin_mse=0.3)),
        "emit": case_emit(),
    }
    for name, arrays in cases.items():
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{name:16s} {os.path.getsize(path) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
