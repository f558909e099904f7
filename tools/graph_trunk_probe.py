"""Probe: how much of the reference (transformers) trunk's training time is host launch overhead?  Times forward +
backward of ModernBertModel (22 layers, A.X-Encoder-base shape, autocast bf16) eagerly and replayed from CUDA graphs
(torch.cuda.make_graphed_callables: same modules, same kernels, same numerics)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_axbase_model(bias_offset=-2.0).to(dev).train()
trunk = model.model.model  # ModernBertModel
g = torch.Generator().manual_seed(1)


def batch(B, S):
    ids = torch.randint(6, 49999, (B, S), generator=g).to(dev)
    mask = torch.ones(B, S, dtype=torch.int64, device=dev)
    return ids, mask


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


out = {}
for B, S in ((64, 256), (64, 64)):
    ids, mask = batch(B, S)

    def eager():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h = trunk(input_ids=ids, attention_mask=mask)[0]
        h.float().square().mean().backward()

    out[f"eager_fwd_bwd_ms_{B}x{S}"] = timed(eager)

    class Wrap(torch.nn.Module):
        def __init__(self, t):
            super().__init__()
            self.t = t

        def forward(self, ids_f, mask_f):
            return self.t(input_ids=ids_f.long(), attention_mask=mask_f.long())[0]

    try:
        w = Wrap(trunk)
        ids_f, mask_f = ids.float(), mask.float()
        with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
            graphed = torch.cuda.make_graphed_callables(w, (ids_f, mask_f))

        def replay():
            with torch.autocast("cuda", dtype=torch.bfloat16, cache_enabled=False):
                h = graphed(ids_f, mask_f)
            h.float().square().mean().backward()

        out[f"graphed_fwd_bwd_ms_{B}x{S}"] = timed(replay)
    except Exception as e:  # noqa: BLE001
        out[f"graphed_error_{B}x{S}"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out), flush=True)
