"""Run under torchrun (>= 2 ranks, one GPU each): SPLADELossV33 with cross-GPU negatives through the NVLink
peer-memory exchange (``exchange=PeerExchange``) against the NCCL all-gather / reduce-scatter statement of the
same step (``gather_negatives=True``) on identical inputs, then time both (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tools/peer_check.py --batch 64 --vocab 50000 --out gpurun_out/peer_check_n2.json

Rank 0 prints ``PEER_CHECK PASS`` (or FAIL with the offending quantity) and one JSON line with the timings.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402


def reps(n, v, seed, density, dev):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, v, generator=g) * (torch.rand(n, v, generator=g) < density)
    return x.to(dev)


def step(fn, q, p, n, autocast, scale):
    q, p, n = (t.clone().requires_grad_(True) for t in (q, p, n))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        loss, d = fn(q, p, n, global_step=20000)
    (loss * scale).backward()
    return loss.detach(), d, q.grad, p.grad, n.grad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--vocab", type=int, default=50000)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--density", type=float, default=0.02)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-timing", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, V, k = a.batch, a.vocab, a.k
    q = reps(B, V, 100 + rank, a.density, dev)
    p = reps(B, V, 200 + rank, a.density, dev)
    n = reps(B * k, V, 300 + rank, a.density, dev).view(B, k, V)
    scale = 1.0 / (1 + rank)

    x = splade_b200.PeerExchange(B, V)
    f_peer = splade_b200.SPLADELossV33(exchange=x)
    f_nccl = splade_b200.SPLADELossV33(gather_negatives=True)

    ok, worst = True, {}
    for autocast in (False, True):
        for it in range(3):  # both window halves, and the first one again
            got = step(f_peer, q, p, n, autocast, scale)
            want = step(f_nccl, q, p, n, autocast, scale)
            x.check()
            for name, g_, w_, tol in (("loss", got[0], want[0], 1e-6), ("dq", got[2], want[2], 1e-5),
                                      ("dp", got[3], want[3], 5e-5), ("dn", got[4], want[4], 1e-5)):
                err = float((g_ - w_).abs().max()) / (float(w_.abs().max()) + 1e-12)
                worst[f"{name}_autocast{int(autocast)}"] = max(worst.get(f"{name}_autocast{int(autocast)}", 0.0), err)
                if not err <= tol:
                    ok = False
            for key in got[1]:
                if abs(got[1][key] - want[1][key]) > 1e-5 * max(1.0, abs(want[1][key])):
                    ok = False
                    worst[f"dict_{key}"] = (got[1][key], want[1][key])
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item() > 0.5)

    timing = {}
    if not a.no_timing:
        for name, fn in (("nccl", f_nccl), ("peer", f_peer)):
            for autocast in (True,):
                for _ in range(5):
                    step(fn, q, p, n, autocast, scale)
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.steps):
                    step(fn, q, p, n, autocast, scale)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                timing[f"{name}_loss_fwd_bwd_ms"] = float(t.item())
        x.check()
        timing["speedup_peer_over_nccl"] = timing["nccl_loss_fwd_bwd_ms"] / timing["peer_loss_fwd_bwd_ms"]

    if rank == 0:
        print("PEER_CHECK", "PASS" if ok else "FAIL", json.dumps(worst))
        line = dict(what="SPLADELossV33 fwd+bwd with cross-GPU negatives: NVLink peer-memory exchange vs NCCL "
                         "all-gather + reduce-scatter (includes the loss_dict D2H sync of every step)",
                    n_gpus=world, batch_per_gpu=B, vocab=V, k=k, density=a.density, steps=a.steps, parity_ok=ok,
                    max_rel_err=worst, **timing)
        print(json.dumps(line))
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "w") as f:
                f.write(json.dumps(line) + "\n")
    x.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
