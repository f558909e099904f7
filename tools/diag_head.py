"""Diagnostic for the fused head kernel on a GPU box: prints mismatch patterns against the oracle for a ladder
of shapes (not a test; used while bringing the tcgen05 path up)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from oracle import splade_oracle as O  # noqa: E402
from tests.gpu_util import make_head_inputs, run_head  # noqa: E402


def one(B, S, H, V, **kw):
    dev = torch.device("cuda:0")
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=B * 1000 + S, device=dev, **kw)
    t0 = time.time()
    rep, arg, tokw = run_head(hidden, W, bias, mask, True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    h, w, b, m = (t.cpu().numpy() for t in (hidden, W, bias, mask))
    rr, ra, rt = O.head_forward(h, w, b, m, "bf16")
    got, ga, gt = rep.cpu().numpy(), arg.cpu().numpy(), tokw.cpu().numpy()
    bad_rep = ~np.isclose(got, rr, rtol=2e-6, atol=1e-7)
    bad_arg = ga != ra
    bad_tok = ~np.isclose(gt, rt, rtol=2e-6, atol=1e-7)
    print(f"B{B} S{S} H{H} V{V} {kw}: {dt*1e3:.1f} ms | rep bad {bad_rep.sum()}/{bad_rep.size} "
          f"argmax bad {bad_arg.sum()} tokw bad {bad_tok.sum()}/{bad_tok.size} "
          f"| max|rep| got {got.max():.4f} ref {rr.max():.4f}", flush=True)
    if bad_rep.any():
        idx = np.argwhere(bad_rep)[:8]
        for b_, v_ in idx:
            print(f"   rep[{b_},{v_}] got {got[b_, v_]:.6f} ref {rr[b_, v_]:.6f} arg got {ga[b_, v_]} ref {ra[b_, v_]}")
        print("   bad per doc:", bad_rep.sum(1)[:16], " bad vocab%128 hist:",
              np.bincount(np.argwhere(bad_rep)[:, 1] % 128, minlength=128).reshape(4, 32).sum(1))
    return bad_rep.sum() == 0


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    ok = True
    for args in [(1, 16, 64, 128), (1, 256, 64, 128), (2, 256, 128, 256), (8, 64, 128, 1000),
                 (3, 7, 64, 130), (5, 100, 64, 777), (2, 300, 64, 260), (4, 256, 768, 2048)]:
        ok &= one(*args)
    print("ALL OK" if ok else "MISMATCHES")
