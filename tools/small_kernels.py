"""Launch each small kernel of the hot path a few times at the BASELINE shapes; meant to be run under
`ncu --metrics gpu__time_duration.sum` so that per-kernel DEVICE time is measured without host launch gaps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402

dev = torch.device("cuda:0")
V = 50000
g = torch.Generator().manual_seed(0)


def reps(n, dens):
    return (torch.rand(n, V, generator=g) * (torch.rand(n, V, generator=g) < dens)).to(dev)


q, p = reps(64, 0.05), reps(64, 0.08)
for k in (1, 7):
    n = reps(64 * k, 0.08).view(64, k, V)
    fn = splade_b200.SPLADELossV33()
    for _ in range(2):
        qq, pp, nn = (t.clone().requires_grad_(True) for t in (q, p, n))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _ = fn(qq, pp, nn if k > 1 else nn[:, 0], global_step=100)
        loss.backward()
x = torch.log1p(torch.relu(torch.randn(512, V, generator=g).to(torch.bfloat16).float())).to(dev)
sp = reps(512, 0.002)
for _ in range(2):
    splade_b200.topk(x, 64, None)
    splade_b200.topk(sp, 64, None)
    splade_b200.topk(x[:64].contiguous(), 100, None)
torch.cuda.synchronize()
print("done")
