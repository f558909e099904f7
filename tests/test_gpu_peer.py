"""-m gpu: cross-GPU negatives through NVLink peer memory (splade_peer_*, peer_exchange.py).

One GPU is enough for the arithmetic, the window layout and the flag barriers: ``PeerExchange.virtual_world`` puts
the W windows on the same device and every virtual rank runs on its own CUDA stream, so the barrier kernels
really wait for each other.  The expected values come from the path that is already pinned to the oracle
(``V33LossFunction`` on the rank-order concatenation, tests/test_gpu_loss.py::test_gathered_negatives_virtual_world).
The CUDA-IPC mapping between processes needs two GPUs (``test_two_processes_ipc``, skipped otherwise).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

os.environ.setdefault("SPLADE_B200_PEER_TIMEOUT_MS", "5000")  # read once by the library, before the first barrier

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reps(n, v, seed, density, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, v, generator=g) * (torch.rand(n, v, generator=g) < density)
    return x.to(dev)


def _expected(q_all, p_all, n_all, W, Bq, k, cfg, scale, t_pos=None, t_neg=None):
    """Per-rank losses / scalars / dq / dnegs and the summed positive gradients from the concatenation path."""
    from splade_b200.ops import V33LossFunction
    losses, scal, dq, dn = [], [], [], []
    dp_sum = torch.zeros_like(p_all)
    for r in range(W):
        sl = slice(r * Bq, (r + 1) * Bq)
        q = q_all[sl].clone().requires_grad_(True)
        docs = p_all.clone().requires_grad_(True)
        n = n_all[sl].clone().requires_grad_(True)
        loss, s = V33LossFunction.apply(q, docs, n, None, None if t_pos is None else t_pos[sl],
                                        None if t_neg is None else t_neg[sl], dict(cfg, label_offset=r * Bq))
        (loss * scale[r]).backward()
        losses.append(float(loss.detach()))
        scal.append(s[:9].cpu().numpy())
        dq.append(q.grad.clone())
        dn.append(n.grad.clone())
        dp_sum += docs.grad
    return losses, scal, dq, dn, dp_sum


def _run_virtual(xs, q_all, p_all, n_all, W, Bq, cfg, scale, backward=True, t_pos=None, t_neg=None):
    from splade_b200.ops import V33PeerLossFunction
    cur = torch.cuda.current_stream()
    streams = [torch.cuda.Stream() for _ in range(W)]
    leaves, outs = [], []
    for r in range(W):
        sl = slice(r * Bq, (r + 1) * Bq)
        streams[r].wait_stream(cur)
        with torch.cuda.stream(streams[r]):
            q = q_all[sl].clone().requires_grad_(True)
            p = p_all[sl].clone().requires_grad_(True)
            n = n_all[sl].clone().requires_grad_(True)
            loss, s = V33PeerLossFunction.apply(q, p, n, None if t_pos is None else t_pos[sl],
                                                None if t_neg is None else t_neg[sl], cfg, xs[r])
            leaves.append((q, p, n))
            outs.append((loss, s))
    if backward:
        for r in range(W):
            with torch.cuda.stream(streams[r]):
                (outs[r][0] * scale[r]).backward()
    for st in streams:
        cur.wait_stream(st)
    torch.cuda.synchronize()
    for x in xs:
        x.check()
    return leaves, outs


@pytest.mark.parametrize("W,Bq,V,k,autocast", [(2, 8, 3000, 1, False), (4, 8, 3000, 2, False),
                                              (4, 8, 3000, 1, True), (2, 70, 1000, 1, True),
                                              (3, 5, 260, 3, False)])
def test_peer_exchange_matches_concatenation(cuda_device, W, Bq, V, k, autocast):
    import splade_b200
    q_all = _reps(W * Bq, V, 21, 0.1, cuda_device)
    p_all = _reps(W * Bq, V, 22, 0.1, cuda_device)
    n_all = _reps(W * Bq * k, V, 23, 0.1, cuda_device).view(W * Bq, k, V)
    cfg = dict(round_bf16=autocast, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3,
               lambda_neg=2e-3, lambda_kd=0.0, lambda_mmse=0.0)
    scale = [1.0 / (1 + r) for r in range(W)]  # a different upstream gradient per rank
    losses, scal, dq, dn, dp_sum = _expected(q_all, p_all, n_all, W, Bq, k, cfg, scale)

    xs = splade_b200.PeerExchange.virtual_world(Bq, V, W, cuda_device)
    try:
        for epoch in range(3):  # both window halves and the reuse of the first one
            leaves, outs = _run_virtual(xs, q_all, p_all, n_all, W, Bq, cfg, scale)
            for r in range(W):
                sl = slice(r * Bq, (r + 1) * Bq)
                assert float(outs[r][0]) == pytest.approx(losses[r], rel=1e-6), (epoch, r)
                np.testing.assert_allclose(outs[r][1][:9].cpu().numpy(), scal[r], rtol=1e-6, atol=1e-7)
                q, p, n = leaves[r]
                torch.testing.assert_close(q.grad, dq[r], rtol=1e-5, atol=1e-8)
                torch.testing.assert_close(n.grad, dn[r], rtol=1e-5, atol=1e-8)
                # autocast: both statements multiply the SAME bf16-rounded G and q (tcgen05 in the concatenation
                # path, fp32 FMAs on rounded operands in the pull-GEMM) - they differ by fp32 summation order only
                torch.testing.assert_close(p.grad, dp_sum[sl], rtol=1e-4 if autocast else 2e-5, atol=1e-7)
    finally:
        for x in xs:
            x.close()


@pytest.mark.parametrize("autocast", [False, True])
def test_peer_exchange_with_margin_mse_teachers(cuda_device, autocast):
    """MarginMSE through the gathered path: the teacher term adds to gin[label] inside the exchange
    window G, i.e. to a column the peers read - loss, loss_dict and all three gradients must equal the
    concatenation statement."""
    import splade_b200
    W, Bq, V, k = 3, 8, 2048, 2
    q_all = _reps(W * Bq, V, 41, 0.1, cuda_device)
    p_all = _reps(W * Bq, V, 42, 0.1, cuda_device)
    n_all = _reps(W * Bq * k, V, 43, 0.1, cuda_device).view(W * Bq, k, V)
    g = torch.Generator().manual_seed(44)
    t_pos = (torch.randn(W * Bq, generator=g) * 3 + 5).to(cuda_device)
    t_neg = (torch.randn(W * Bq, k, generator=g) * 3).to(cuda_device)
    cfg = dict(round_bf16=autocast, temperature=1.0, kd_temperature=1.0, lambda_q=1e-2, lambda_d=3e-3,
               lambda_neg=2e-3, lambda_kd=0.0, lambda_mmse=0.7)
    scale = [1.0 / (1 + r) for r in range(W)]
    losses, scal, dq, dn, dp_sum = _expected(q_all, p_all, n_all, W, Bq, k, cfg, scale, t_pos, t_neg)
    assert all(s_[6] > 0 for s_ in scal)          # the MarginMSE term is really there
    xs = splade_b200.PeerExchange.virtual_world(Bq, V, W, cuda_device)
    try:
        for epoch in range(2):
            leaves, outs = _run_virtual(xs, q_all, p_all, n_all, W, Bq, cfg, scale, t_pos=t_pos, t_neg=t_neg)
            for r in range(W):
                sl = slice(r * Bq, (r + 1) * Bq)
                assert float(outs[r][0]) == pytest.approx(losses[r], rel=1e-6), (epoch, r)
                np.testing.assert_allclose(outs[r][1][:9].cpu().numpy(), scal[r], rtol=1e-6, atol=1e-7)
                q, p, n = leaves[r]
                torch.testing.assert_close(q.grad, dq[r], rtol=1e-5, atol=1e-8)
                torch.testing.assert_close(n.grad, dn[r], rtol=1e-5, atol=1e-8)
                torch.testing.assert_close(p.grad, dp_sum[sl], rtol=1e-4 if autocast else 2e-5, atol=1e-7)
    finally:
        for x in xs:
            x.close()


def test_forward_only_epochs_do_not_block(cuda_device):
    """Evaluation (no backward) advances only the forward flags; a later training step still lines up."""
    import splade_b200
    W, Bq, V, k = 2, 8, 1024, 1
    q_all, p_all = _reps(W * Bq, V, 31, 0.1, cuda_device), _reps(W * Bq, V, 32, 0.1, cuda_device)
    n_all = _reps(W * Bq, V, 33, 0.1, cuda_device).view(W * Bq, k, V)
    cfg = dict(round_bf16=False, temperature=1.0, kd_temperature=1.0, lambda_q=0.0, lambda_d=0.0, lambda_neg=0.0,
               lambda_kd=0.0, lambda_mmse=0.0)
    scale = [1.0] * W
    losses, _, _, _, dp_sum = _expected(q_all, p_all, n_all, W, Bq, k, cfg, scale)
    xs = splade_b200.PeerExchange.virtual_world(Bq, V, W, cuda_device)
    try:
        for _ in range(3):
            _, outs = _run_virtual(xs, q_all, p_all, n_all, W, Bq, cfg, scale, backward=False)
            assert [float(o[0]) for o in outs] == pytest.approx(losses, rel=1e-6)
        leaves, _ = _run_virtual(xs, q_all, p_all, n_all, W, Bq, cfg, scale)
        for r in range(W):
            torch.testing.assert_close(leaves[r][1].grad, dp_sum[r * Bq:(r + 1) * Bq], rtol=2e-5, atol=1e-7)
    finally:
        for x in xs:
            x.close()


def test_missing_peer_times_out_instead_of_hanging(cuda_device):
    """Only rank 0 of a 2-rank virtual world shows up: the barrier gives up after the timeout and check() raises."""
    import splade_b200
    from splade_b200.ops import V33PeerLossFunction
    Bq, V = 4, 512
    xs = splade_b200.PeerExchange.virtual_world(Bq, V, 2, cuda_device)
    try:
        q, p = _reps(Bq, V, 1, 0.2, cuda_device), _reps(Bq, V, 2, 0.2, cuda_device)
        n = _reps(Bq, V, 3, 0.2, cuda_device).view(Bq, 1, V)
        cfg = dict(round_bf16=False, temperature=1.0, kd_temperature=1.0, lambda_q=0.0, lambda_d=0.0,
                   lambda_neg=0.0, lambda_kd=0.0, lambda_mmse=0.0)
        V33PeerLossFunction.apply(q, p, n, None, None, cfg, xs[0])
        with pytest.raises(RuntimeError, match="waited in vain for rank 1"):
            xs[0].check()
    finally:
        for x in xs:
            x.close()


def test_peer_exchange_argument_errors(cuda_device):
    import splade_b200
    from splade_b200.ops import V33PeerLossFunction
    with pytest.raises(RuntimeError, match="process group"):
        splade_b200.PeerExchange(8, 512)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        splade_b200.PeerExchange.virtual_world(8, 510, 2, cuda_device)
    xs = splade_b200.PeerExchange.virtual_world(8, 512, 2, cuda_device)
    try:
        q = _reps(4, 512, 1, 0.2, cuda_device)  # batch 4, windows were sized for 8
        cfg = dict(round_bf16=False, temperature=1.0, kd_temperature=1.0, lambda_q=0.0, lambda_d=0.0,
                   lambda_neg=0.0, lambda_kd=0.0, lambda_mmse=0.0)
        with pytest.raises(RuntimeError, match="built for batch 8"):
            V33PeerLossFunction.apply(q, q, q.view(4, 1, 512), None, None, cfg, xs[0])
        fn = splade_b200.SPLADELossV33(lambda_kd=1.0, exchange=xs[0])
        q8 = _reps(8, 512, 1, 0.2, cuda_device)
        with pytest.raises(RuntimeError, match="KL-KD"):
            fn(q8, q8, q8, teacher_scores=torch.zeros(8, 8, device=cuda_device))
    finally:
        for x in xs:
            x.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="CUDA IPC between ranks needs two GPUs")
def test_two_processes_ipc():
    """Two real ranks: windows mapped with CUDA IPC, SPLADELossV33(exchange=...) against the NCCL
    all-gather / reduce-scatter path on the same inputs (tools/peer_check.py prints PEER_CHECK PASS)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "peer_check.py"), "--batch", "16",
           "--vocab", "4096", "--steps", "3", "--no-timing"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "PEER_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
