"""Full encode (ids -> sparse vectors) at the A.X-Encoder-base shape: native packed trunk + fused head vs the
reference trunk (transformers under autocast) + fused head vs the reference's whole chain.  One JSON line per case."""
import json, os, sys, tempfile, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from oracle import torch_reference as TR
from transformers import AutoModelForMaskedLM, ModernBertConfig
dev = torch.device("cuda:0")
V, H = 50000, 768
cfg = ModernBertConfig(vocab_size=V, hidden_size=H, intermediate_size=1152, num_hidden_layers=22,
                       num_attention_heads=12, pad_token_id=V - 1, bos_token_id=0, eos_token_id=1, cls_token_id=0,
                       sep_token_id=1, max_position_embeddings=16384, global_attn_every_n_layers=3, local_attention=128)
torch.manual_seed(0)
hf = AutoModelForMaskedLM.from_config(cfg)
with tempfile.TemporaryDirectory() as d:
    hf.save_pretrained(d)
    ours = splade_b200.SPLADEModernBERT(model_name=d).to(dev).eval()
with torch.no_grad():
    ours.model.decoder.bias.normal_(0.0, 0.02).sub_(float(os.environ.get("BIAS_OFFSET", "0")))
trunk = splade_b200.NativeTrunk(ours.model)
ref = TR.ReferenceSpladeModule(ours.model).eval()
g = torch.Generator().manual_seed(1)


def timed(fn, n):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, B, S, lo, hi in (("docs 64x256 full", 64, 256, 256, 256), ("docs 64x256 U[32,256]", 64, 256, 32, 256),
                           ("queries 512x64 U[4,24]", 512, 64, 4, 24)):
    lens = torch.randint(lo, hi + 1, (B,), generator=g)
    ids = torch.randint(6, V - 1, (B, S), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64)
    ids = torch.where(mask.bool(), ids, torch.full_like(ids, V - 1)).to(dev); mask = mask.to(dev)
    row = {"case": name, "real_tokens": int(lens.sum())}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a = trunk.encode(ids, mask); b = ours.encode(ids, mask)
        cos = torch.nn.functional.cosine_similarity(a, b, dim=1)
        row["cosine_min_vs_reference_trunk"] = float(cos.min())
        row["max_abs_diff"] = float((a - b).abs().max())
        top_a, top_b = a.topk(64, dim=1).indices, b.topk(64, dim=1).indices
        row["top64_overlap"] = float(sum(len(set(x.tolist()) & set(y.tolist())) for x, y in zip(top_a, top_b)) / (64 * B))
        if name.startswith("docs 64x256 full"):
            # the reference's inference path is fp32 without autocast (benchmark/encoders.py:309-318)
            with torch.autocast("cuda", enabled=False):
                c = ref(ids, mask)[0]
            row["cosine_min_vs_fp32_reference_chain"] = float(torch.nn.functional.cosine_similarity(a, c, dim=1).min())
            row["cosine_min_transformers_bf16_vs_fp32_reference_chain"] = float(
                torch.nn.functional.cosine_similarity(b, c, dim=1).min())
            top_c = c.topk(64, dim=1).indices
            row["top64_overlap_vs_fp32_reference_chain"] = float(
                sum(len(set(x.tolist()) & set(y.tolist())) for x, y in zip(top_a, top_c)) / (64 * B))
            row["top64_overlap_transformers_bf16_vs_fp32"] = float(
                sum(len(set(x.tolist()) & set(y.tolist())) for x, y in zip(top_b, top_c)) / (64 * B))
            del c
        row["native_ms"] = timed(lambda: trunk.encode(ids, mask), 10)
        row["reference_trunk_plus_fused_head_ms"] = timed(lambda: ours.encode(ids, mask), 5)
        row["reference_chain_ms"] = timed(lambda: ref(ids, mask)[0], 3)
    row["docs_per_s_native"] = B / row["native_ms"] * 1e3
    row["speedup_vs_reference_chain"] = row["reference_chain_ms"] / row["native_ms"]
    print(json.dumps(row), flush=True)
