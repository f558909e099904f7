"""The fused compute+exchange kernels of the NVLink peer exchange under ncu on ONE GPU: W virtual ranks whose windows
all live on this device (PeerExchange.virtual_world), one CUDA stream per rank, so the flag barriers really wait for each
other and `peer_gather_kernel` / `peer_ddocs_kernel` / `peer_bwd_sync_kernel` move exactly the bytes they would pull over
NVLink - (W-1) * B * V * 4 per rank in each direction - out of the other windows (here: HBM of the same GPU)."""
import os
import sys

os.environ.setdefault("SPLADE_B200_PEER_TIMEOUT_MS", "1000")  # read once by the library

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from splade_b200.ops import V33PeerLossFunction  # noqa: E402

dev = torch.device("cuda:0")
W, Bq, V = int(os.environ.get("W", "4")), 64, 50000
g = torch.Generator().manual_seed(0)


def reps(n):
    return (torch.rand(n, V, generator=g) * (torch.rand(n, V, generator=g) < 0.006)).to(dev)


xs = splade_b200.PeerExchange.virtual_world(Bq, V, W, dev)
cfg = dict(round_bf16=True, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3, lambda_neg=3e-3,
           lambda_kd=0.0, lambda_mmse=0.0)
streams = [torch.cuda.Stream() for _ in range(W)]
cur = torch.cuda.current_stream()
for _ in range(int(os.environ.get("REPS", "2"))):
    outs = []
    for r in range(W):
        streams[r].wait_stream(cur)
        with torch.cuda.stream(streams[r]):
            q, p, n = reps(Bq).requires_grad_(True), reps(Bq).requires_grad_(True), reps(Bq).view(Bq, 1, V).requires_grad_(True)
            outs.append(V33PeerLossFunction.apply(q, p, n, None, None, cfg, xs[r])[0])
    for r in range(W):
        with torch.cuda.stream(streams[r]):
            outs[r].backward()
    for st in streams:
        cur.wait_stream(st)
    torch.cuda.synchronize()
for x in xs:
    try:  # under ncu every launch is serialised, so a flag barrier cannot be answered and gives up (status word set):
        x.check()  # the exchange kernels still move exactly the same bytes, which is what the capture is for
    except RuntimeError as e:
        print("note:", str(e)[:120])
    x.close()
print("ok", W * Bq * V * 4 * (W - 1) // W, "bytes pulled per rank per direction")
