"""A few launches of the packed (variable-length) head kernel on a query-like batch (for `ncu --set full`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from tests.gpu_util import make_head_inputs, query_like_mask
dev = torch.device("cuda:0")
case = os.environ.get("CASE", "queries")
B, S, lo, hi = (512, 64, 4, 24) if case == "queries" else (64, 256, 32, 256)
H, V = 768, 50000
hidden, W, bias, _ = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False)
h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
mask = query_like_mask(B, S, lo, hi, seed=9, device=dev)
for _ in range(int(os.environ.get("N", "4"))):
    splade_b200.head_forward(h16, w16, bias, mask, os.environ.get("TOKW", "0") == "1", pack_mode=1)
torch.cuda.synchronize(); print("done", int(mask.sum()))
