"""Quick CUDA-event timing of the fused head at the headline shapes (bring-up aid, not the bench)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from oracle import torch_reference as TR  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
for (B, S) in [(64, 256), (512, 64)]:
    H, V = 768, 50000
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    for tokw in (False, True):
        for _ in range(3):
            splade_b200.head_forward(h16, w16, bias, mask, tokw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            splade_b200.head_forward(h16, w16, bias, mask, tokw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        tf = 2.0 * B * S * H * V / ms / 1e9
        print(f"B{B} S{S} tokw={tokw}: {ms:.3f} ms  {tf:.0f} TFLOP/s  {B / ms * 1e3:.0f} docs/s", flush=True)
    with torch.no_grad():
        for _ in range(2):
            TR.head_chain_autocast(hidden, W, bias, mask)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            TR.head_chain_autocast(hidden, W, bias, mask)
        e1.record()
        torch.cuda.synchronize()
        print(f"B{B} S{S} torch autocast chain: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
