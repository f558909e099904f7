"""One launch of every secondary kernel at the BASELINE shapes (for `ncu --set full`)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from tests.gpu_util import make_head_inputs
dev = torch.device("cuda:0")
B, S, H, V = 64, 256, 768, 50000
hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False)
h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, True)
g = torch.randn(B, V, device=dev)
splade_b200.head_backward(g, rep, arg, h16, w16)
gq = torch.Generator().manual_seed(0)
r = lambda n, d: (torch.rand(n, V, generator=gq) * (torch.rand(n, V, generator=gq) < d)).to(dev).requires_grad_(True)
q, p, n = r(64, 0.05), r(64, 0.08), r(64, 0.08)
with torch.autocast("cuda", dtype=torch.bfloat16):
    loss, _ = splade_b200.SPLADELossV33()(q, p, n, global_step=100)
loss.backward()
x = torch.log1p(torch.relu(torch.randn(512, V, generator=gq).to(torch.bfloat16).float())).to(dev)
splade_b200.topk(x, 64, None)
torch.cuda.synchronize(); print("done")
