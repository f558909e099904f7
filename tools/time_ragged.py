"""Head forward on right-padded (ragged) batches: time vs the full-length batch of the same shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from tests.gpu_util import make_head_inputs
dev = torch.device("cuda:0")
B, S, H, V = 64, 256, 768, 50000
hidden, W, bias, _ = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False)
h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
g = torch.Generator().manual_seed(3)
cases = {"full": torch.full((B,), S), "U[32,256]": torch.randint(32, S + 1, (B,), generator=g),
         "U[16,128]+one full": torch.cat([torch.tensor([S]), torch.randint(16, 129, (B - 1,), generator=g)])}
for rep in range(2):
    for name, lens in cases.items():
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(dev)
        for _ in range(5): splade_b200.head_forward(h16, w16, bias, mask, False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30): splade_b200.head_forward(h16, w16, bias, mask, False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        toks = int(lens.sum())
        print(f"{name:22s} real tokens {toks:6d} ({100*toks/(B*S):5.1f}%)  {ms:.3f} ms  {B/ms*1e3:8.0f} docs/s  "
              f"{2.0*toks*H*V/ms/1e9:6.0f} TFLOP/s on real tokens", flush=True)
