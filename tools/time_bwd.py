"""Head backward of ONE training micro-step (query 64x64, positive 64x256, negatives 64k x 256 on the tied decoder
matrix; train_v33_ddp.py:339-343, 364): three single-segment calls + the two dense adds autograd performs on their
[V,H] gradients  vs  the multi-segment sweep (splade_head_bwd_multi) and the two-phase form autograd drives.
Dense (random init) and trained-like sparse regimes.  Bring-up / optimisation aid; judged numbers come from bench.py."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
H, V = 768, 50000
K_NEG = int(os.environ.get("K_NEG", "1"))


COLD = os.environ.get("COLD", "1") == "1"   # flush L2 before every timed call (the default) or leave it warm


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if COLD:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


out = {}
for off, tag in ((0.0, "dense"), (-2.0, "sparse")):
    _, W, bias, _ = make_head_inputs(1, 8, H, V, seed=1, device=dev, bias_offset=off)
    w16 = W.to(torch.bfloat16).contiguous()
    segs = []
    for i, (B, S) in enumerate(((64, 64), (64, 256), (64 * K_NEG, 256))):
        hidden, _, _, mask = make_head_inputs(B, S, H, V, seed=2 + i, device=dev, ragged=False)
        h16 = hidden.to(torch.bfloat16).contiguous()
        rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, False)
        segs.append((torch.randn(B, V, device=dev) * (rep > 0), rep, arg, h16, splade_b200.rep_bitmap(rep)))
    nnz = [int((s[1] > 0).sum()) // s[1].shape[0] for s in segs]

    def three_singles():
        r = [splade_b200.head_backward(*s[:4], w16) for s in segs]
        dW = r[0][1] + r[1][1]          # what autograd's accumulation does with three dense gradients
        dW = dW + r[2][1]
        return dW

    def multi():
        return splade_b200.head_backward_multi(segs, w16)

    def two_phase():
        ph = [splade_b200.head_backward_dhidden(*s[:4], w16, active=s[4]) for s in segs]
        return splade_b200.head_backward_dw(segs, H, V)

    def phase1_only():
        return [splade_b200.head_backward_dhidden(*s[:4], w16, active=s[4]) for s in segs]

    def dw_only():
        return splade_b200.head_backward_dw(segs, H, V)

    def bitmaps_only():   # paid in the forward (rep is L2-hot there; here it is cold: an upper bound)
        return [splade_b200.rep_bitmap(s[1]) for s in segs]

    row = {"nnz_per_doc": nnz, "k_neg": K_NEG, "l2": "flushed before every call" if COLD else "warm"}
    for name, fn in (("three_single_calls_plus_autograd_adds_ms", three_singles), ("multi_ms", multi),
                     ("two_phase_ms", two_phase), ("phase1_bucket_dhidden_ms", phase1_only), ("dw_sweep_ms", dw_only),
                     ("rep_bitmaps_cold_ms", bitmaps_only)):
        row[name] = timeit(fn)
    docs = sum(s[1].shape[0] for s in segs)
    rows = sum(s[3].shape[0] * s[3].shape[1] for s in segs)
    # compulsory HBM bytes of the multi form: g/rep/argmax at the active terms (one 32 B sector each), the bitmaps,
    # dW + dbias and every dHidden written once
    row["algorithmic_bytes"] = sum(nnz) * 64 * 3 * 32 + docs * (V // 8) + V * H * 4 + V * 4 + rows * H * 4
    row["multi_GBps"] = row["algorithmic_bytes"] / row["multi_ms"] / 1e6
    out[tag] = row
    print(tag, json.dumps(row), flush=True)
