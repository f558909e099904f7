"""CUDA-event timings of every kernel on the hot path at the BASELINE shapes, with the HBM roofline of each
(bring-up / optimisation aid; the judged numbers come from bench.py)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from tests.gpu_util import make_head_inputs  # noqa: E402

dev = torch.device("cuda:0")
HBM = 6548.2e9
flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, n=10, l2_flush=True):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if l2_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def report(name, ms, bytes_=None, flops=None):
    s = f"{name:58s} {ms * 1e3:9.1f} us"
    if bytes_:
        s += f"  {bytes_ / ms / 1e6:8.0f} GB/s ({bytes_ / ms / 1e6 / (HBM / 1e9) * 100:5.1f}% of HBM copy peak)"
    if flops:
        s += f"  {flops / ms / 1e9:8.0f} TFLOP/s"
    print(s, flush=True)


B, S, H, V = 64, 256, 768, 50000
for bias_offset, tag in ((0.0, "dense (random init)"), (-2.0, "sparse (bias -2)")):
    hidden, W, bias, mask = make_head_inputs(B, S, H, V, seed=1, device=dev, ragged=False, bias_offset=bias_offset)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    rep, arg, _ = splade_b200.head_forward(h16, w16, bias, mask, False)
    nnz = int((rep > 0).sum())
    g = torch.randn(B, V, device=dev)
    ms = timeit(lambda: splade_b200.head_backward(g, rep, arg, h16, w16))
    # compulsory HBM bytes: dW + dbias + d_hidden written, g/rep/argmax read, W and hidden read once
    comp = V * H * 4 + V * 4 + B * S * H * 4 + 3 * B * V * 4 + V * H * 2 + B * S * H * 2
    gathered = 2 * nnz * H * 2
    report(f"head_bwd {tag}, nnz/doc {nnz // B}", ms, comp)
    print(f"{'':58s} gathered rows (L2): {gathered / 1e9:.2f} GB -> {gathered / ms / 1e9:.2f} TB/s")
    if bias_offset == 0.0:
        ms = timeit(lambda: splade_b200.head_forward(h16, w16, bias, mask, False))
        report("head_fwd C2 (no token_weights)", ms, flops=2.0 * B * S * H * V)
        ms = timeit(lambda: splade_b200.head_forward(h16, w16, bias, mask, True))
        report("head_fwd C2 (+ token_weights)", ms, flops=2.0 * B * S * H * V)

for N in (64, 448):
    rep = torch.rand(N, V, device=dev) * (torch.rand(N, V, device=dev) < 0.1)
    ms = timeit(lambda: splade_b200.flops_nnz(rep))
    report(f"flops_nnz [{N},{V}]", ms, N * V * 4)
q = torch.rand(64, V, device=dev) * (torch.rand(64, V, device=dev) < 0.1)
for N in (64, 512):
    docs = torch.rand(N, V, device=dev) * (torch.rand(N, V, device=dev) < 0.1)
    for rb in (False, True):
        ms = timeit(lambda: splade_b200.scores(q, docs, rb))
        report(f"scores [64,{V}]x[{N},{V}] round_bf16={rb}", ms, (64 + N) * V * 4, 2.0 * 64 * N * V)
for k in (1, 7):
    negs = torch.rand(64, k, V, device=dev)
    ms = timeit(lambda: splade_b200.rowdot(q, negs))
    report(f"rowdot k={k}", ms, (64 + 64 * k) * V * 4)

fn = splade_b200.SPLADELossV33()
for k in (1, 7):
    qq = q.clone().requires_grad_(True)
    pp = (torch.rand(64, V, device=dev) * (torch.rand(64, V, device=dev) < 0.1)).requires_grad_(True)
    nn_ = (torch.rand(64, k, V, device=dev) * (torch.rand(64, k, V, device=dev) < 0.1)).requires_grad_(True)

    def fwd():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return fn(qq, pp, nn_ if k > 1 else nn_[:, 0], global_step=100)[0]

    def fwdbwd():
        fwd().backward()

    ms = timeit(fwd)
    report(f"SPLADELossV33 forward k={k} (incl. the one D2H sync)", ms, (2 + k) * 64 * V * 4 * 2)
    ms2 = timeit(fwdbwd)
    report(f"SPLADELossV33 forward+backward k={k}", ms2, (2 + k) * 64 * V * 4 * 4)
    os.environ["SPLADE_B200_LOSS_BWD_TC"] = "0"
    ms3 = timeit(fwdbwd)
    os.environ["SPLADE_B200_LOSS_BWD_TC"] = "1"
    report(f"   ... with the fp32 CUDA-core backward (round 1)", ms3, (2 + k) * 64 * V * 4 * 4)

# gathered negatives as at 8 GPUs (N = 512 in-batch documents), one rank's view, virtual (no exchange)
from splade_b200.ops import V33LossFunction  # noqa: E402
for N in (64, 512):
    qq = q.clone().requires_grad_(True)
    dd = (torch.rand(N, V, device=dev) * (torch.rand(N, V, device=dev) < 0.1)).requires_grad_(True)
    nn_ = (torch.rand(64, 1, V, device=dev) * (torch.rand(64, 1, V, device=dev) < 0.1)).requires_grad_(True)
    cfg = dict(round_bf16=True, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3, lambda_neg=3e-3,
               lambda_kd=0.0, lambda_mmse=0.0, label_offset=0)
    for tc in ("1", "0"):
        os.environ["SPLADE_B200_LOSS_BWD_TC"] = tc
        loss, _ = V33LossFunction.apply(qq, dd, nn_, None, None, None, cfg)

        def bwd():
            torch.autograd.grad(loss, [qq, dd, nn_], retain_graph=True)

        ms = timeit(bwd)
        report(f"loss backward only, N={N} docs, {'tcgen05' if tc == '1' else 'fp32 CUDA cores'}", ms,
               (64 * 3 + N * 2) * V * 4)
    os.environ["SPLADE_B200_LOSS_BWD_TC"] = "1"

for (Bt, k) in ((64, 64), (512, 64), (512, 100)):
    x = torch.log1p(torch.relu(torch.randn(Bt, V, device=dev).to(torch.bfloat16).float()))
    ms = timeit(lambda: splade_b200.topk(x, k, None))
    report(f"topk B={Bt} k={k}", ms, Bt * V * 4)
