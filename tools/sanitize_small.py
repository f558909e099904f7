"""Tiny invocation of every kernel (for compute-sanitizer --tool memcheck on the GPU box)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
for (B, S, H, V) in [(3, 37, 64, 300), (2, 300, 64, 200)]:
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=dev)
    h = hidden.clone().requires_grad_(True)
    w = W.clone().requires_grad_(True)
    b = bias.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        rep, tokw = splade_b200.splade_head(h, w, b, mask)
        n = rep.roll(1, 0)
        loss, d = splade_b200.SPLADELossV33(lambda_margin_mse=0.1, lambda_kd=0.1)(
            rep, rep.flip(0), torch.stack([n, rep], 1), global_step=3,
            teacher_scores=torch.randn(B, B, device=dev), teacher_pos_scores=torch.rand(B, device=dev),
            teacher_neg_scores=torch.rand(B, 2, device=dev))
    loss.backward()
    ids, vals, counts = splade_b200.topk(rep.detach(), 5, None)
    torch.cuda.synchronize()
    print(B, S, H, V, float(loss), counts.tolist())
print("sanitize run ok")
