"""Forward + backward of the native training trunk (NativeTrunkTrain) vs the transformers trunk at the A.X-Encoder-base
shape, trunk only, autocast bf16; with PROFILE=1 a torch.profiler table of the native path's kernels."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import splade_b200  # noqa: E402

dev = torch.device("cuda:0")
model = bench.build_axbase_model(bias_offset=-2.0).to(dev).train()
hf = model.model.model
native = splade_b200.NativeTrunkTrain(model.model)
g = torch.Generator().manual_seed(1)


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


NCU = os.environ.get("NCU") == "1"   # under ncu: one shape, one warm-up, one measured pass
for B, S, lo in (((64, 256, 256),) if NCU else ((64, 256, 256), (64, 64, 64), (64, 256, 32), (64, 64, 4))):
    # lo < S: right-padded batch with U[lo, S] real tokens per document (what the reference's collator produces)
    lens = torch.randint(lo, S + 1, (B,), generator=g)
    lens[0] = S
    mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int64).to(dev)
    ids = torch.randint(6, 49999, (B, S), generator=g).to(dev)

    def run_native(backward=True):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h = native(ids, mask)
        if backward:
            h.float().square().mean().backward()

    def run_hf():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            h = hf(input_ids=ids, attention_mask=mask)[0]
        h.float().square().mean().backward()

    if NCU:
        run_native()
        torch.cuda.synchronize()
        run_native()
        torch.cuda.synchronize()
        continue
    row = {"shape": f"{B}x{S}", "real_tokens": int(lens.sum()), "native_fwd_ms": timed(lambda: run_native(False)), "native_fwd_bwd_ms": timed(run_native),
           "transformers_fwd_bwd_ms": timed(run_hf)}
    print(json.dumps(row), flush=True)
    if os.environ.get("PROFILE") == "1" and S == 256:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            run_native()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
