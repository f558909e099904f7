"""Diagnostic: in-batch score matrix of the reference under autocast vs the scores kernel's bf16 emulation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from oracle import torch_reference as TR  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (B, V, k, dens) in [(4, 5000, 2, 1.0), (64, 50000, 1, 0.05), (8, 50000, 1, 1.0)]:
    q = (torch.rand(B, V, generator=g) * (torch.rand(B, V, generator=g) < dens)).to(dev)
    p = (torch.rand(B, V, generator=g) * (torch.rand(B, V, generator=g) < dens)).to(dev)
    n = (torch.rand(B, k, V, generator=g) * (torch.rand(B, k, V, generator=g) < dens)).to(dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ref_scores = torch.mm(q, p.t())
    raw = splade_b200.scores(q, p, True)
    ours = raw.to(torch.bfloat16)
    exact = (q.double() @ p.double().t())
    exact_bf = (q.to(torch.bfloat16).double() @ p.to(torch.bfloat16).double().t())
    print(f"B{B} V{V}: ref dtype {ref_scores.dtype}; mismatching bf16 scores {(ours != ref_scores).sum().item()} of {B*B}")
    print("  ref  row0", ref_scores[0, :4].float().tolist())
    print("  ours row0", ours[0, :4].float().tolist())
    print("  raw  row0", raw[0, :4].tolist())
    print("  exact(bf16 inputs) row0", exact_bf[0, :4].tolist())
    print("  exact(fp32 inputs) row0", exact[0, :4].tolist())
    fn = splade_b200.SPLADELossV33()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l_ours, d = fn(q, p, n if k > 1 else n[:, 0], global_step=0)
        lam = [fn._lambda_schedule(0, x) for x in (fn.lambda_q, fn.lambda_d, fn.lambda_neg)]
        l_ref = TR.v33_loss(q, p, n if k > 1 else n[:, 0], *lam)
        in_batch = torch.mm(q, p.t()) / 1.0
        hard = (q.unsqueeze(1) * n).sum(-1)
        nce_ref = torch.nn.functional.cross_entropy(torch.cat([in_batch, hard], 1), torch.arange(B, device=dev))
    print(f"  loss ours {float(l_ours):.6f} ref {float(l_ref):.6f} | infonce ours {d['infonce']:.6f} ref {float(nce_ref):.6f}"
          f" | flops q/d/n ours {d['flops_q']:.4f} {d['flops_d']:.4f} {d['flops_neg']:.4f} ref "
          f"{float(TR.flops_reg(q)):.4f} {float(TR.flops_reg(p)):.4f} {float(TR.flops_reg(n.reshape(-1, V))):.4f}")
    print("  hard ours", splade_b200.rowdot(q, n)[0].tolist(), "ref", hard[0].tolist())
