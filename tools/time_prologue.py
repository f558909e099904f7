"""CUDA-event timing of the head prologue (dense -> GELU -> LayerNorm -> bf16) per call: the transformers module +
cast (what the default path ran before) vs the fused prologue, forward and forward+backward, at the C4 shapes."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402,F401
from splade_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
H, eps = 768, 1e-5


def timed(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for T in (16384, 4096):
    g = torch.Generator().manual_seed(T)
    x = torch.randn(T, H, generator=g).to(dev).requires_grad_(True)
    dense_w = (torch.randn(H, H, generator=g) * H ** -0.5).to(dev).requires_grad_(True)
    norm_w = torch.ones(H, device=dev, requires_grad=True)
    d_hidden = torch.randn(T, H, generator=g).to(dev)
    w16 = dense_w.detach().to(torch.bfloat16).contiguous()
    w16_t = w16.t().contiguous()
    out = {"T": T, "H": H}

    def module_fwd():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h = F.layer_norm(F.gelu(F.linear(x, dense_w)), (H,), norm_w, None, eps)
        return h, h.detach().to(torch.bfloat16)

    def module_fwd_bwd():
        h, _ = module_fwd()
        torch.autograd.grad(h, (x, dense_w, norm_w), d_hidden)

    def fused_fwd():
        x16 = x.detach().to(torch.bfloat16)
        return x16, ops.head_prologue_forward(x16, w16, norm_w.detach(), eps, T)

    def fused_fwd_bwd():
        x16, (y, hidden, stats) = fused_fwd()
        dy, dx, dn = ops.head_prologue_backward(d_hidden, y, stats, norm_w.detach(), w16_t, T)
        dx.to(torch.float32)
        ops._mm_f32(dy.t(), x16)

    with torch.no_grad():
        out["module_fwd_ms"] = timed(lambda: module_fwd())
    out["module_fwd_bwd_ms"] = timed(module_fwd_bwd)
    out["fused_fwd_ms"] = timed(fused_fwd)
    out["fused_fwd_bwd_ms"] = timed(fused_fwd_bwd)
    print(json.dumps(out), flush=True)
