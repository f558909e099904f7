"""A/B timing inside ONE process/box: single-CTA vs CTA-pair head kernel (the env switch is read once per process,
so each variant runs in a subprocess; variants alternate to cancel drift)."""
import os, subprocess, sys
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import splade_b200
from tests.gpu_util import make_head_inputs
dev = torch.device("cuda:0")
out = []
for (B, S) in [(64, 256), (512, 64)]:
    hidden, W, bias, mask = make_head_inputs(B, S, 768, 50000, seed=1, device=dev, ragged=False)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    for tokw in (False, True):
        for _ in range(5): splade_b200.head_forward(h16, w16, bias, mask, tokw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): splade_b200.head_forward(h16, w16, bias, mask, tokw)
        e1.record(); torch.cuda.synchronize()
        out.append("%.3f" % (e0.elapsed_time(e1) / 50))
print(" ".join(out))
'''
for rep in range(3):
    for pair in ("0", "1"):
        env = dict(os.environ, SPLADE_B200_CTA_PAIR=pair)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print("pair=" + pair, "C2 / C2+tokw / C3 / C3+tokw ms:", r.stdout.strip() or r.stderr[-300:], flush=True)
