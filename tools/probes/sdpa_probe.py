"""Probe: cost of torch SDPA forward + backward at the trunk's attention shape (64 docs x 12 heads x 256 x 64, bf16),
with the sliding-window / padding mask the reference trunk passes and without any mask."""
import json
import time

import torch
import torch.nn.functional as F

dev = torch.device("cuda:0")
B, nh, S, D = 64, 12, 256, 64
q, k, v = (torch.randn(B, nh, S, D, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
do = torch.randn(B, nh, S, D, device=dev, dtype=torch.bfloat16)
idx = torch.arange(S, device=dev)
win = ((idx[:, None] - idx[None, :]).abs() <= 64)[None, None].expand(B, 1, S, S).contiguous()
addm = torch.zeros(B, 1, S, S, device=dev, dtype=torch.bfloat16).masked_fill(~win, float("-inf"))


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def run(mask):
    def f():
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=0.125)
        torch.autograd.grad(o, (q, k, v), do)
    return f


print(json.dumps({"sdpa_fwd_bwd_ms_no_mask": timed(run(None)), "sdpa_fwd_bwd_ms_bool_window_mask": timed(run(win)),
                  "sdpa_fwd_bwd_ms_additive_bf16_mask": timed(run(addm))}))
