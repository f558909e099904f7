"""SPLADELossV33 forward + backward under autocast at the training shape, local negatives (N = 64 in-batch documents)
and gathered-style (N = 512), for ncu: every kernel of the fused forward and the tcgen05 backward once."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from splade_b200.ops import V33LossFunction  # noqa: E402

dev = torch.device("cuda:0")
V = 50000
g = torch.Generator().manual_seed(0)


def reps(n, dens):
    return (torch.rand(n, V, generator=g) * (torch.rand(n, V, generator=g) < dens)).to(dev)


cfg = dict(round_bf16=True, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3, lambda_neg=3e-3,
           lambda_kd=0.0, lambda_mmse=0.0, label_offset=0)
for N in (64, 512):
    for _ in range(int(os.environ.get("REPS", "2"))):
        q, d, n = reps(64, 0.006).requires_grad_(True), reps(N, 0.006).requires_grad_(True), \
            reps(64, 0.006).view(64, 1, V).requires_grad_(True)
        loss, _ = V33LossFunction.apply(q, d, n, None, None, None, cfg)
        loss.backward()
torch.cuda.synchronize()
print("ok")
