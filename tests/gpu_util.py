"""Helpers shared by the -m gpu tests (the product is called through the C-ABI via splade_b200.ops)."""
import numpy as np
import torch


def bf16r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def make_head_inputs(B, S, H, V, seed, device, ragged=True, bias_offset=0.0, holes=False):
    """Synthetic inputs per SURVEY.md §8(d): hidden ~ N(0,1), W ~ N(0, 0.0176^2) (scaled to keep the logit
    std at small H), bias ~ N(0, 0.02) + offset; lengths ~ U[S/8, S], right padded.  All bf16-representable."""
    g = torch.Generator().manual_seed(seed)
    hidden = bf16r(torch.randn(B, S, H, generator=g))
    W = bf16r(torch.randn(V, H, generator=g) * 0.0176 * (768 / H) ** 0.5)
    bias = bf16r(torch.randn(V, generator=g) * 0.02 + bias_offset)
    if ragged:
        lengths = torch.randint(max(1, S // 8), S + 1, (B,), generator=g)
        lengths[0] = S
    else:
        lengths = torch.full((B,), S)
    mask = (torch.arange(S)[None, :] < lengths[:, None]).to(torch.int32)
    if holes:
        mask[:, min(3, S - 1)] = 0
    return hidden.to(device), W.to(device), bias.to(device), mask.to(device)


def run_head(hidden, W, bias, mask, want_tokw=True):
    import splade_b200
    return splade_b200.head_forward(hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous(),
                                    bias.float().contiguous(), mask.to(torch.int32).contiguous(), want_tokw)


def assert_argmax_parity(got_arg, ref_arg, near_tie, what=""):
    """Bit-exact argmax, except where the reference's top two positions are within one bf16 ulp (a different
    fp32 accumulation order may round the logit the other way there)."""
    got_arg = np.asarray(got_arg).astype(np.int64)
    ref_arg = np.asarray(ref_arg).astype(np.int64)
    bad = got_arg != ref_arg
    n_bad = int(bad.sum())
    n_unexplained = int((bad & ~np.asarray(near_tie)).sum())
    assert n_unexplained == 0, f"{what}: {n_unexplained} argmax mismatches away from any near-tie ({n_bad} total)"
    assert n_bad <= max(2, int(2e-4 * bad.size)), f"{what}: too many near-tie flips: {n_bad} of {bad.size}"
    return n_bad
