"""A few native-trunk encodes at the A.X-Encoder-base shape (for ncu launch lists / captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from transformers import AutoModelForMaskedLM, ModernBertConfig
dev = torch.device("cuda:0")
V, H = 50000, 768
cfg = ModernBertConfig(vocab_size=V, hidden_size=H, intermediate_size=1152, num_hidden_layers=22,
                       num_attention_heads=12, pad_token_id=V - 1, bos_token_id=0, eos_token_id=1, cls_token_id=0,
                       sep_token_id=1, max_position_embeddings=16384, global_attn_every_n_layers=3, local_attention=128)
torch.manual_seed(0)
mlm = AutoModelForMaskedLM.from_config(cfg).to(dev).eval()
trunk = splade_b200.NativeTrunk(mlm)
B, S = int(os.environ.get("B", "64")), int(os.environ.get("S", "256"))
g = torch.Generator().manual_seed(1)
ids = torch.randint(6, V - 1, (B, S), generator=g).to(dev)
lo = int(os.environ.get("LO", str(S)))
lens = torch.randint(lo, S + 1, (B,), generator=g)
mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64).to(dev)
for _ in range(int(os.environ.get("N", "2"))):
    trunk.encode(ids, mask)
torch.cuda.synchronize(); print("done")
