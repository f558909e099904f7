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
