"""One multi-segment head backward of a micro-step (q 64x64, pos 64x256, neg 64x256) for ncu: REGIME=sparse|dense."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
H, V = 768, 50000
off = -2.0 if os.environ.get("REGIME", "sparse") == "sparse" else 0.0
_, W, bias, _ = make_head_inputs(1, 8, H, V, seed=1, device=dev, bias_offset=off)
w16 = W.to(torch.bfloat16).contiguous()
segs = []
for i, (B, S) in enumerate(((64, 64), (64, 256), (64, 256))):
    hidden, _, _, mask = make_head_inputs(B, S, H, V, seed=2 + i, device=dev, ragged=False)
    h16 = hidden.to(torch.bfloat16).contiguous()
    rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, False)
    segs.append((torch.randn(B, V, device=dev) * (rep > 0), rep, arg, h16, splade_b200.rep_bitmap(rep)))
for _ in range(int(os.environ.get("REPS", "3"))):
    splade_b200.head_backward_multi(segs, w16)
torch.cuda.synchronize()
print("ok")
