"""Hot-loop timings of the trunk building blocks at the A.X-Encoder-base shapes (T = 16384 packed tokens)."""
import json, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200
from splade_b200 import _lib as L
dev = torch.device("cuda:0")
T, H, I = int(os.environ.get("T", "16384")), 768, 1152
g = torch.Generator().manual_seed(0)
bf = lambda *s: (torch.randn(*s, generator=g) / math.sqrt(s[-1])).to(torch.bfloat16).to(dev)


def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


A768, A1152 = bf(T, H), bf(T, I)
Wqkv, Wo, Wi, Wo2 = bf(3 * H, H), bf(H, H), bf(2 * I, H), bf(H, I)
x = torch.zeros(T, H, device=dev)
o_qkv = torch.empty(T, 3 * H, dtype=torch.bfloat16, device=dev)
o_h = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
o_i = torch.empty(T, I, dtype=torch.bfloat16, device=dev)
cases = {
    "qkv  bf16  [T,768]x[2304,768]": (lambda: splade_b200.trunk_gemm(L.EPI_BF16, A768, Wqkv, out=o_qkv), 2 * T * H * 3 * H),
    "wo   resid [T,768]x[768,768]": (lambda: splade_b200.trunk_gemm(L.EPI_RESID, A768, Wo, out=x), 2 * T * H * H),
    "wi   geglu [T,768]x[2304,768]": (lambda: splade_b200.trunk_gemm(L.EPI_GEGLU, A768, Wi, out=o_i), 2 * T * H * 2 * I),
    "wo2  resid [T,1152]x[768,1152]": (lambda: splade_b200.trunk_gemm(L.EPI_RESID, A1152, Wo2, out=x), 2 * T * I * H),
    "head gelu  [T,768]x[768,768]": (lambda: splade_b200.trunk_gemm(L.EPI_GELU, A768, Wo, out=o_h), 2 * T * H * H),
}
if os.environ.get("ONLY_ATTN") == "1":
    cases = {}
for name, (fn, flop) in cases.items():
    us = timed(fn)
    print(json.dumps({"kernel": name, "us": round(us, 1), "tflops": round(flop / us / 1e6, 1)}), flush=True)
# the same GEMMs through cuBLAS (torch.matmul, bf16) for reference
for name, a, w in () if os.environ.get("ONLY_ATTN") == "1" else (("cublas qkv", A768, Wqkv), ("cublas wo", A768, Wo), ("cublas wi", A768, Wi), ("cublas wo2", A1152, Wo2)):
    us = timed(lambda: torch.matmul(a, w.t()))
    print(json.dumps({"kernel": name, "us": round(us, 1), "tflops": round(2 * a.shape[0] * a.shape[1] * w.shape[0] / us / 1e6, 1)}), flush=True)
# attention: 64 documents x 256 tokens, 12 heads
B, S, nh = T // 256, 256, 12
mask = torch.ones(B, S, dtype=torch.int32, device=dev)
qkv = bf(T, 3 * H)
for window in (-1, 64):
    us = timed(lambda: splade_b200.trunk_attention(qkv, mask, nh, window), 20)
    print(json.dumps({"kernel": f"attention window={window} (includes the layout prepass)", "us": round(us, 1)}), flush=True)
