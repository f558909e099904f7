import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from tests.gpu_util import make_head_inputs
dev = torch.device("cuda:0")
B, S, H, V = 64, 256, 768, 50000
for off in (0.0, -2.0):
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False, bias_offset=off)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, False)
    g = torch.randn(B, V, device=dev)
    for _ in range(2):
        splade_b200.head_backward(g, rep, arg, h16, w16)
torch.cuda.synchronize(); print("done")
