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


def run_head(hidden, W, bias, mask, want_tokw=True, pack_mode=None, plan_out=None):
    import splade_b200
    return splade_b200.head_forward(hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous(),
                                    bias.float().contiguous(), mask.to(torch.int32).contiguous(), want_tokw,
                                    pack_mode=pack_mode, plan_out=plan_out)


def query_like_mask(B, S, lo, hi, seed, device):
    """Right-padded batch whose documents hold U[lo, hi] real tokens (queries: a handful of tokens in S=64)."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(lo, hi + 1, (B,), generator=g)
    return (torch.arange(S)[None, :] < lengths[:, None]).to(torch.int32).to(device)


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


def assert_weights_parity(got, ref, what="", max_flip_frac=2e-4):
    """Pooled weights: the SAME bf16 logit gives log1p equal to an fp32 ulp or two.  A different fp32
    accumulation order (tensor-core tree vs BLAS) may land on the other side of a bf16 rounding boundary for
    ~1e-5 of the cells: those must stay within ONE bf16 ulp (2^-7 relative on
    the logit) and be rare.  Everything is inside the north_star tolerance (1e-2 relative)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    exact = np.isclose(got, ref, rtol=2e-6, atol=1e-7)
    n_flip = int((~exact).sum())
    assert n_flip <= max(2, int(max_flip_frac * exact.size)), f"{what}: {n_flip} of {exact.size} weights differ"
    if n_flip:
        lg, lr = np.expm1(got[~exact]), np.expm1(ref[~exact])   # back to the logit domain
        one_ulp = np.maximum(np.abs(lr), 2.0 ** -126) * 2.0 ** -7 * 1.01
        assert np.all(np.abs(lg - lr) <= one_ulp), f"{what}: a weight differs by more than one bf16 ulp"
    np.testing.assert_allclose(got, ref, rtol=1e-2, atol=1e-2, err_msg=what)
    return n_flip
