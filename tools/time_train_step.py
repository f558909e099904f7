"""Full V33 training micro-step at the BASELINE config-4 shape, ours vs the reference statement (1..8 GPUs).

The step is the reference's (src/train/cli/train_v33_ddp.py:337-374, restated in splade_b200.training):
three encoder calls (64 queries x 64, 64 positives x 256, 64 negatives x 256 tokens) + SPLADELossV33 under
autocast(bf16), backward, clip 1.0, AdamW - on a random-init A.X-Encoder-base-shaped ModernBERT (22 layers,
hidden 768, vocab 50 000, 149 M parameters).  Both sides run the SAME transformers trunk with autograd:

  ours       splade_b200.SPLADEModernBERT + splade_b200.SPLADELossV33 (fused head fwd/bwd + loss kernels)
  reference  oracle/torch_reference.py: HF logits -> relu -> log1p -> mask -> max, loss with stock torch ops

Prints one JSON line (CUDA-event time per step, tokens/s, peak memory).  decoder.bias is offset so the pooled
vectors are sparse, as for a trained model.

Under torchrun (WORLD_SIZE > 1) both sides are wrapped in DistributedDataParallel exactly as the reference does
(src/train/cli/train_v33_ddp.py:539-544), every rank gets its own batches (seed 42 + rank, :503), times are the
max over ranks and tokens/s is the whole job's; ``--negatives peer|nccl`` additionally turns on cross-GPU
negatives for OUR side (the reference has local negatives only):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
        tools/time_train_step.py --negatives peer --out gpurun_out/train_step_n8.json
"""
import argparse
import copy
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from oracle import torch_reference as TR  # noqa: E402
from splade_b200.training import StepConfig, synthetic_triplet_batch, train_micro_step  # noqa: E402

VOCAB, HIDDEN = 50000, 768


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--layers", type=int, default=22)
    ap.add_argument("--bias-offset", type=float, default=-2.0)
    ap.add_argument("--negatives", default="local", choices=["local", "nccl", "peer"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP
        dist.init_process_group("nccl", device_id=dev)
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=VOCAB, hidden_size=HIDDEN, intermediate_size=1152, num_hidden_layers=a.layers,
                           num_attention_heads=12, pad_token_id=VOCAB - 1, bos_token_id=0, eos_token_id=1,
                           cls_token_id=0, sep_token_id=1, max_position_embeddings=16384,
                           global_attn_every_n_layers=3, local_attention=128)
    torch.manual_seed(0)
    hf = AutoModelForMaskedLM.from_config(cfg)
    with torch.no_grad():
        hf.decoder.bias.copy_(torch.randn(VOCAB, generator=torch.Generator().manual_seed(2)) * 0.02 + a.bias_offset)
    with tempfile.TemporaryDirectory() as d:
        hf.save_pretrained(d)
        ours = splade_b200.SPLADEModernBERT(model_name=d)
    del hf
    ref = TR.ReferenceSpladeModule(copy.deepcopy(ours.model))
    B = a.batch
    batches = [synthetic_triplet_batch(B, 64, 256, 1, VOCAB, VOCAB - 1, seed=1000 * (42 + rank) + i, ragged=False)
               for i in range(a.steps + a.warmup)]
    tokens = world * B * (64 + 256 + 256)
    out = {"what": "full V33 training micro-step: trunk (transformers, autograd) + head + SPLADELossV33 + clip + "
                   "AdamW under autocast(bf16); DDP when n_gpus > 1",
           "n_gpus": world, "negatives": a.negatives if world > 1 else "local", "batch_per_gpu": B,
           "tokens_per_step": tokens, "layers": a.layers, "steps": a.steps, "warmup": a.warmup,
           "params": sum(p.numel() for p in ours.parameters()), "bias_offset": a.bias_offset}
    sched = splade_b200.SPLADELossV33()

    def ref_loss(anchor_repr, positive_repr, negative_repr, global_step, **kw):
        lam = [sched._lambda_schedule(global_step, x) for x in (sched.lambda_q, sched.lambda_d, sched.lambda_neg)]
        loss = TR.v33_loss(anchor_repr, positive_repr, negative_repr, *lam)
        return loss, {"nonzero_d": float((positive_repr > 0).sum(1).float().mean())}  # one sync, like .item() calls

    exchange = None
    ours_loss = splade_b200.SPLADELossV33()
    if world > 1 and a.negatives == "peer":
        exchange = splade_b200.PeerExchange(B, VOCAB)
        ours_loss = splade_b200.SPLADELossV33(exchange=exchange)
    elif world > 1 and a.negatives == "nccl":
        ours_loss = splade_b200.SPLADELossV33(gather_negatives=True)

    for tag, model, loss_fn in (("ours", ours, ours_loss), ("reference", ref, ref_loss)):
        opt = None
        try:
            model = model.to(dev).train()
            if world > 1:
                model = DDP(model, device_ids=[local], find_unused_parameters=False)
            opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.01)
            scfg = StepConfig(gradient_accumulation_steps=1)
            gs, losses, d = 20000, [], {}
            torch.cuda.reset_peak_memory_stats(dev)
            for i in range(a.warmup):
                loss, d, gs = train_micro_step(model, loss_fn, opt, None, batches[i], i, gs, scfg, dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.warmup, a.warmup + a.steps):
                loss, d, gs = train_micro_step(model, loss_fn, opt, None, batches[i], i, gs, scfg, dev)
                losses.append(loss.detach())
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            if world > 1:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            out[tag] = {"ms_per_step": ms, "tokens_per_s": tokens / ms * 1e3,
                        "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30,
                        "last_loss": float(losses[-1]), "nonzero_d": d.get("nonzero_d")}
        except Exception as e:  # noqa: BLE001 - keep the other side's number
            out[tag] = {"error": f"{type(e).__name__}: {e}"[:300]}
        (model.module if world > 1 and hasattr(model, "module") else model).to("cpu")
        del opt, model
        torch.cuda.empty_cache()
    if "ms_per_step" in out.get("ours", {}) and "ms_per_step" in out.get("reference", {}):
        out["speedup"] = out["reference"]["ms_per_step"] / out["ours"]["ms_per_step"]
    if exchange is not None:
        exchange.check()
        exchange.close()
    line = json.dumps(out)
    if rank == 0:
        print(line)
    if world > 1:
        dist.destroy_process_group()
    if a.out and rank == 0:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
