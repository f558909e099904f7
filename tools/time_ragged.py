"""Head forward on right-padded (ragged) batches: padded tiling vs packed (variable-length) tiling vs the
device-side auto plan, at the document shape (C2: 64 x 256) and the query shape (C3: 512 x 64).
Prints one JSON line per case (bring-up / profiling aid; bench.py carries the judged numbers)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
H, V = 768, 50000
ITERS = int(os.environ.get("ITERS", "30"))


def lens_cases(B, S, g):
    if S == 256:
        return {"full": torch.full((B,), S),
                "U[32,256]": torch.randint(32, S + 1, (B,), generator=g),
                "U[16,128]+one full": torch.cat([torch.tensor([S]), torch.randint(16, 129, (B - 1,), generator=g)])}
    return {"full": torch.full((B,), S),
            "U[4,24] (queries)": torch.randint(4, 25, (B,), generator=g),
            "U[8,64]": torch.randint(8, S + 1, (B,), generator=g)}


def timed(fn, iters=ITERS):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for (B, S) in ((64, 256), (512, 64)):
    hidden, W, bias, _ = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    g = torch.Generator().manual_seed(3)
    for name, lens in lens_cases(B, S, g).items():
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(dev)
        toks = int(lens.sum())
        row = {"shape": f"{B}x{S}", "lengths": name, "real_tokens": toks, "real_frac": round(toks / (B * S), 4)}
        for tokw in (False, True):
            for label, mode in (("padded", 0), ("packed", 1), ("auto", -1)):
                plan = torch.zeros(4, dtype=torch.int32, device=dev)
                ms = timed(lambda: splade_b200.head_forward(h16, w16, bias, mask, tokw, pack_mode=mode, plan_out=plan))
                key = label + ("_tokw" if tokw else "")
                row[key + "_ms"] = round(ms, 4)
                if label == "auto":
                    row["auto_picked" + ("_tokw" if tokw else "")] = "packed" if int(plan[0]) == 1 else "padded"
        best = min(row["padded_ms"], row["packed_ms"])
        row["tflops_on_real_tokens_best"] = round(2.0 * toks * H * V / best / 1e9, 1)
        row["speedup_packed_over_padded"] = round(row["padded_ms"] / row["packed_ms"], 3)
        print(json.dumps(row), flush=True)
