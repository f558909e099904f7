"""-m gpu: error behaviour of the boundary (every failure is a RuntimeError carrying splade_last_error(), never a
fallback or a crash) and run-to-run determinism of the forward paths."""
import pytest
import torch

from tests.gpu_util import make_head_inputs, run_head

pytestmark = pytest.mark.gpu


def test_unsupported_shapes_and_arguments_raise(cuda_device):
    import splade_b200
    dev = cuda_device
    h = torch.zeros(2, 8, 96, dtype=torch.bfloat16, device=dev)            # H = 96 is not a multiple of 64
    w = torch.zeros(32, 96, dtype=torch.bfloat16, device=dev)
    with pytest.raises(RuntimeError, match="multiple of 64"):
        splade_b200.head_forward(h, w, torch.zeros(32, device=dev), torch.ones(2, 8, dtype=torch.int32, device=dev))
    with pytest.raises(RuntimeError, match=r"k=2000 outside"):
        splade_b200.topk(torch.rand(2, 4096, device=dev), 2000)
    with pytest.raises(AssertionError):                                     # dtype contract of the raw op
        splade_b200.head_forward(h.float(), w, torch.zeros(32, device=dev),
                                 torch.ones(2, 8, dtype=torch.int32, device=dev))
    lib = splade_b200.load_library()
    assert lib.splade_head_fwd(None, None, None, None, 1, 1, 64, 1, None, None, None, None, 0, None) == 1
    assert b"null pointer" in lib.splade_last_error()
    # KL-KD with gathered negatives is rejected by the host mirror, MarginMSE without teachers is simply off
    fn = splade_b200.SPLADELossV33(lambda_margin_mse=0.5)
    q = torch.rand(4, 64, device=dev)
    loss, d = fn(q, q.flip(0), q.roll(1, 0))
    assert d["margin_mse"] == 0.0 and torch.isfinite(loss)
    # the library still works after errors
    hidden, W, bias, mask = make_head_inputs(2, 16, 64, 100, seed=0, device=dev)
    rep, _, _ = run_head(hidden, W, bias, mask)
    assert torch.isfinite(rep).all()


def test_forward_paths_are_deterministic(cuda_device):
    import splade_b200
    hidden, W, bias, mask = make_head_inputs(16, 128, 256, 20000, seed=5, device=cuda_device)
    a = run_head(hidden, W, bias, mask)
    b = run_head(hidden, W, bias, mask)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    rep = a[0]
    t1 = splade_b200.topk(rep, 64, None)
    t2 = splade_b200.topk(rep, 64, None)
    assert all(torch.equal(x, y) for x, y in zip(t1, t2))
    fn = splade_b200.SPLADELossV33()
    outs = []
    for _ in range(2):
        q, p, n = (rep[:8].clone().requires_grad_(True), rep[8:].clone().requires_grad_(True),
                   rep[8:].flip(0).clone().requires_grad_(True))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss, _ = fn(q, p, n, global_step=7)
        loss.backward()
        outs.append((loss.detach(), q.grad, p.grad, n.grad))
    assert all(torch.equal(x, y) for x, y in zip(*outs))
    # the head backward is bit-reproducible too: dW / dbias sum over the batch in a fixed order,
    g = torch.randn_like(rep)
    h16, w16 = hidden.to(torch.bfloat16).contiguous(), W.to(torch.bfloat16).contiguous()
    b1 = splade_b200.head_backward(g, rep, a[1], h16, w16)
    b2 = splade_b200.head_backward(g, rep, a[1], h16, w16)
    assert torch.equal(b1[1], b2[1]) and torch.equal(b1[2], b2[2])
    assert torch.equal(b1[0], b2[0])   # dHidden: buckets list their terms in ascending term id (ordered counting sort)
