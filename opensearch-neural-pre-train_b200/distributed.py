"""Multi-GPU plumbing for the hot path: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
box, gloo in CPU tests).

* Encode (BASELINE configs 2, 3, 5) shards documents across ranks with NO data-path collective
  (``shard_range``); only the final top-k buffers are concatenated in rank order.
* Training (config 4) keeps the reference's semantics by default (per-rank local negatives, DDP gradient
  all-reduce, src/train/cli/train_v33_ddp.py:539-544).  ``all_gather_documents`` is the one real exchange
  step of the north-star extension: positives are all-gathered forward and their gradients are
  reduce-scattered (sum) backward, so that, after DDP's gradient averaging, the result equals the gradient of
  the mean InfoNCE over all W*B queries.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of ceil(n/W) items for `rank` (last ranks may be short or empty)."""
    per = (n_items + world_size - 1) // world_size
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


class _AllGatherDocs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, local: torch.Tensor, group):
        rank, world = world_info(group)
        ctx.group, ctx.rank, ctx.world = group, rank, world
        local = local.contiguous()
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out

    @staticmethod
    def backward(ctx, g_all: torch.Tensor):
        g_all = g_all.contiguous()
        n = g_all.shape[0] // ctx.world
        g_local = torch.empty((n,) + tuple(g_all.shape[1:]), dtype=g_all.dtype, device=g_all.device)
        if g_all.is_cuda:
            dist.reduce_scatter_tensor(g_local, g_all, op=dist.ReduceOp.SUM, group=ctx.group)
        else:  # gloo has no reduce_scatter: all_reduce then slice (CPU tests only)
            dist.all_reduce(g_all, op=dist.ReduceOp.SUM, group=ctx.group)
            g_local.copy_(g_all[ctx.rank * n:(ctx.rank + 1) * n])
        return g_local, None


def all_gather_documents(local_docs: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """[B,V] per rank -> ([W*B,V] in rank order, label offset rank*B).  Identity when not distributed."""
    rank, world = world_info(group)
    if world == 1:
        return local_docs, 0
    return _AllGatherDocs.apply(local_docs, group), rank * local_docs.shape[0]


def bind_to_device_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPU cores NVML reports as local to `device_index` (and therefore its
    later pinned-host allocations to that socket's memory).  With 8 ranks streaming host buffers at ~40 GB/s
    each, cross-socket traffic is what limits end-to-end scaling.  Returns False when NVML is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False
