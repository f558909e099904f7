"""CPU, gloo, world_size 2: the N>1 host logic (document all-gather with reduce-scatter backward, shard
ranges, bulk-encode sharding).  The device kernels are not involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = globals()[fn_name](rank, world)
    finally:
        dist.destroy_process_group()


def _spawn(fn_name, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn_name, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _gather_case(rank, world):
    import splade_b200
    B, V = 3, 5
    local = (torch.arange(B * V, dtype=torch.float32).view(B, V) + 100 * rank).requires_grad_(True)
    docs, off = splade_b200.all_gather_documents(local)
    assert off == rank * B and docs.shape == (world * B, V)
    # every rank weights the gathered rows differently: d local = sum over ranks of their weights on MY rows
    w = torch.arange(world * B, dtype=torch.float32).view(-1, 1) * (rank + 1)
    (docs * w).sum().backward()
    return docs.detach().clone(), local.grad.clone()


def test_all_gather_documents_forward_and_reduce_scatter_backward():
    out = _spawn("_gather_case", 2)
    B, V, world = 3, 5, 2
    want = torch.cat([torch.arange(B * V, dtype=torch.float32).view(B, V) + 100 * r for r in range(world)])
    for r, (docs, grad) in enumerate(out):
        assert torch.equal(docs, want)                                           # rank order
        rows = torch.arange(r * B, (r + 1) * B, dtype=torch.float32).view(-1, 1)
        assert torch.equal(grad, (rows * sum(k + 1 for k in range(world))).expand(B, V))


def _bulk_case(rank, world):
    import splade_b200
    N, S, V, k = 11, 4, 16, 3

    class FakeModel(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def encode(self, ids, mask):
            return torch.nn.functional.one_hot(ids[:, 0] % V, V).float() * (ids[:, :1].float() + 1)

    class FakeEmitter:
        def select(self, rep, top_k):
            vals, ids = rep.topk(top_k, dim=1)
            return ids.int(), vals, (vals > 0).sum(1).int()

    ids = torch.arange(N).view(N, 1).repeat(1, S)
    out = splade_b200.bulk_encode(FakeModel(), ids, torch.ones_like(ids), k, emitter=FakeEmitter(), batch_size=2,
                                  autocast_dtype=None)
    return out[0][:, 0].tolist(), out[3]


def test_bulk_encode_shards_contiguously_without_collectives():
    out = _spawn("_bulk_case", 2)
    assert out[0][1] == (0, 6) and out[1][1] == (6, 11)
    assert out[0][0] + out[1][0] == [i % 16 for i in range(11)]                  # rank-order concatenation


def _bf16_hook_case(rank, world):
    from torch.nn.parallel import DistributedDataParallel as DDP

    from splade_b200.training import enable_bf16_gradient_compression
    torch.manual_seed(0)
    model = torch.nn.Linear(16, 8)
    ddp = DDP(model)
    enable_bf16_gradient_compression(ddp)
    x = torch.full((4, 16), float(rank + 1))
    ddp(x).sum().backward()
    return model.weight.grad.clone(), model.bias.grad.clone()


def test_bf16_gradient_compression_hook_averages_like_ddp():
    """N3 (opt-in): gradients are averaged over ranks as DDP does, within bf16 rounding of the exact mean."""
    try:
        out = _spawn("_bf16_hook_case", 2)
    except Exception as e:  # a gloo build without bf16 all-reduce: the hook is NCCL-only there
        pytest.skip(f"gloo cannot all-reduce bf16 here: {type(e).__name__}")
    # d(sum(Wx+b))/dW = sum over the batch of x = 4 * (rank+1) per input column; mean over ranks = 6; bias: 4
    for w_grad, b_grad in out:
        torch.testing.assert_close(w_grad, torch.full((8, 16), 6.0), rtol=2 ** -8, atol=0)
        torch.testing.assert_close(b_grad, torch.full((8,), 4.0), rtol=2 ** -8, atol=0)
