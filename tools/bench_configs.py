"""Measure the remaining BASELINE.json configurations (3, 4-variant, 5) on the GPU box and write one JSON file.

    python tools/bench_configs.py [--docs-per-gpu 8192] [--out gpurun_out/configs.json]
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py ...

  config 3  query encode, batch 512 x L=64 with top-k pruning (k=64, real blocklist size): head-only and full model
  config 4b k=7 hard negatives at per-GPU batch 16 (configs/train_v34_multi_neg.yaml) hot-path training step
  config 5  bulk corpus encoding sharded across the ranks: full model (A.X-Encoder-base shape, random init) ->
            fused head -> GPU top-k -> compact (ids, weights) buffers on the host; docs/s per job

All timings are CUDA events / wall clock around device-synchronised regions, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402

H, V = 768, 50000


def cuda_time(fn, n, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def build_model(dev):
    from transformers import AutoModelForMaskedLM, ModernBertConfig
    cfg = ModernBertConfig(vocab_size=V, hidden_size=H, intermediate_size=1152, num_hidden_layers=22,
                           num_attention_heads=12, pad_token_id=V - 1, bos_token_id=0, eos_token_id=1,
                           cls_token_id=0, sep_token_id=1, max_position_embeddings=16384,
                           global_attn_every_n_layers=3, local_attention=128)
    torch.manual_seed(0)
    hf = AutoModelForMaskedLM.from_config(cfg)
    with torch.no_grad():  # sparse-ish regime so that top-k sees realistic candidate counts
        hf.decoder.bias.copy_(torch.randn(V, generator=torch.Generator().manual_seed(2)) * 0.02 - 2.0)
    with tempfile.TemporaryDirectory() as d:
        hf.save_pretrained(d)
        return splade_b200.SPLADEModernBERT(model_name=d).to(dev).eval()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs-per-gpu", type=int, default=8192)
    ap.add_argument("--out", default="gpurun_out/configs.json")
    ap.add_argument("--native-trunk", action="store_true", help="config 5 / config 3 full model through the native trunk")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        splade_b200.distributed.bind_to_device_numa_node(local)
        dist.init_process_group("nccl", device_id=dev)

    def max_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {"n_gpus": world}
    g = torch.Generator().manual_seed(7 + rank)
    block = torch.zeros(V, dtype=torch.uint8)
    block[list(range(37)) + [64, 95, V - 1]] = 1
    emitter = splade_b200.SparseEmitter(block.to(dev))

    # ---------------- config 3: query encode 512 x 64 + top-k (head only, hidden resident)
    Bq, Sq, k = 512, 64, 64
    hidden = torch.randn(Bq, Sq, H, generator=g).to(torch.bfloat16).to(dev)
    W = (torch.randn(V, H, generator=g) * 0.0176).to(torch.bfloat16).to(dev)
    bias = (torch.randn(V, generator=g) * 0.02).to(dev)
    mask = torch.ones(Bq, Sq, dtype=torch.int32, device=dev)

    def c3_head():
        rep, _, _ = splade_b200.head_forward(hidden, W, bias, mask, want_token_weights=False)
        return emitter.select(rep, k)

    ms = max_ranks(cuda_time(c3_head, 20))
    ms_head = max_ranks(cuda_time(lambda: splade_b200.head_forward(hidden, W, bias, mask, False), 20))
    out["config3_head_topk"] = {"batch": Bq, "seq_len": Sq, "k": k, "ms": ms, "ms_head_only": ms_head,
                                "ms_topk": ms - ms_head, "queries_per_s": world * Bq / ms * 1e3,
                                "head_tflops": 2.0 * Bq * Sq * H * V / ms_head / 1e9}
    del hidden

    # ---------------- config 4b: k = 7 negatives, per-GPU batch 16 (train_v34_multi_neg.yaml), hot path only
    B4, k7 = 16, 7
    hq = torch.randn(B4, 64, H, generator=g).to(dev).requires_grad_(True)
    hp = torch.randn(B4, 256, H, generator=g).to(dev).requires_grad_(True)
    hn = torch.randn(B4 * k7, 256, H, generator=g).to(dev).requires_grad_(True)
    Wf = (torch.randn(V, H, generator=g) * 0.0176).to(dev).requires_grad_(True)
    bf = (torch.randn(V, generator=g) * 0.02 - 1.0).to(dev).requires_grad_(True)
    loss_fn = splade_b200.SPLADELossV33(lambda_neg=0.01)
    ones = lambda n, s: torch.ones(n, s, dtype=torch.int64, device=dev)  # noqa: E731

    def c4b():
        for t in (hq, hp, hn, Wf, bf):
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a, _ = splade_b200.splade_head(hq, Wf, bf, ones(B4, 64))
            p, _ = splade_b200.splade_head(hp, Wf, bf, ones(B4, 256))
            n, _ = splade_b200.splade_head(hn, Wf, bf, ones(B4 * k7, 256))
            loss, _ = loss_fn(a, p, n.view(B4, k7, -1), global_step=5000)
        loss.backward()
        if world > 1:
            dist.all_reduce(Wf.grad, op=dist.ReduceOp.AVG)

    ms = max_ranks(cuda_time(c4b, 10))
    toks = B4 * (64 + 256 + 256 * k7)
    out["config4b_k7_batch16"] = {"ms_per_step": ms, "tokens_per_s": world * toks / ms * 1e3,
                                  "tokens_per_step_per_gpu": toks}
    del hq, hp, hn, Wf, bf

    # ---------------- config 5: bulk encode, full model, sharded, top-k emit
    model = build_model(dev)
    if args.native_trunk:
        model.use_native_trunk(True)   # packed-token trunk + head in one C-ABI call
    n_local = args.docs_per_gpu
    n_total = n_local * world
    ids_all = torch.randint(6, V - 1, (n_local, 256), generator=g)       # this rank's shard of the id stream
    mask_all = torch.ones(n_local, 256, dtype=torch.int64)
    ids_all, mask_all = ids_all.pin_memory(), mask_all.pin_memory()
    # warm-up
    splade_b200.bulk_encode(model, ids_all[:128], mask_all[:128], 64, emitter=emitter, batch_size=64, group=None)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # every rank encodes ITS shard (the corpus tensor passed here already is the shard -> run as world 1)
    lo, hi = 0, n_local
    out_ids, out_vals, out_counts = [], [], []
    with torch.no_grad():
        for b0 in range(lo, hi, 64):
            ib = ids_all[b0:b0 + 64].to(dev, non_blocking=True)
            mb = mask_all[b0:b0 + 64].to(dev, non_blocking=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                rep = model.encode(ib, mb)
            i, v, c = emitter.select(rep, 64)
            out_ids.append(i.to("cpu", non_blocking=True))
            out_vals.append(v.to("cpu", non_blocking=True))
            out_counts.append(c.to("cpu", non_blocking=True))
    torch.cuda.synchronize()
    dt = max_ranks(time.perf_counter() - t0)
    nnz = float(torch.cat(out_counts).float().mean())
    # ---- the emit leg behind it (benchmark/indexer.py:112-138 / scripts/neural_sparse_search_aws.py:980-1036): the
    #      compact top-k buffers of this rank's shard -> the OpenSearch _bulk request body (splade_emit_bulk, host C),
    #      (a) alone on one host thread, (b) on a side thread WHILE the GPU encodes the shard again (ctypes releases
    #      the GIL), i.e. what a pipelined indexer pays in wall time.
    import threading
    shard = (torch.cat(out_ids), torch.cat(out_vals), torch.cat(out_counts))
    doc_ids = [f"d{rank}_{i}" for i in range(n_local)]
    chunk = 8192

    def emit_all():
        total = 0
        for c0 in range(0, n_local, chunk):
            body = emitter.to_bulk(shard[0][c0:c0 + chunk], shard[1][c0:c0 + chunk], shard[2][c0:c0 + chunk],
                                   "bench", doc_ids[c0:c0 + chunk], round_decimals=4)
            total += len(body)
        return total

    from splade_b200._lib import load_library
    lib = load_library()
    emit_all()  # warm-up (token table, allocator)
    prev_threads = lib.splade_emit_set_threads(1)
    t0 = time.perf_counter()
    body_bytes = emit_all()
    dt_emit_one = time.perf_counter() - t0
    lib.splade_emit_set_threads(prev_threads)   # automatic: the host's cores, at most 16, document ranges per thread
    t0 = time.perf_counter()
    body_bytes = emit_all()
    dt_emit = time.perf_counter() - t0
    box = {}
    th = threading.Thread(target=lambda: box.setdefault("bytes", emit_all()))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th.start()
    with torch.no_grad():
        for b0 in range(lo, hi, 64):
            ib = ids_all[b0:b0 + 64].to(dev, non_blocking=True)
            mb = mask_all[b0:b0 + 64].to(dev, non_blocking=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                rep = model.encode(ib, mb)
            emitter.select(rep, 64)
    torch.cuda.synchronize()
    dt_enc_overlapped = time.perf_counter() - t0
    th.join()
    dt_both = max_ranks(time.perf_counter() - t0)
    emit = {"what": "SparseEmitter.to_bulk (splade_emit_bulk): _bulk action + source lines, sparse_vector format, "
                    "4-decimal weights, for this rank's shard; exact integer number formatting, document ranges on the "
                    "emitter's worker threads (automatic count)",
            "host_cores": os.cpu_count(), "seconds_alone_one_thread": dt_emit_one,
            "MB_per_s_alone_one_thread": body_bytes / dt_emit_one / 1e6,
            "docs": n_local, "body_bytes": body_bytes, "bytes_per_doc": body_bytes / max(1, n_local),
            "seconds_alone": dt_emit, "docs_per_s_alone": n_local / dt_emit, "MB_per_s_alone": body_bytes / dt_emit / 1e6,
            "seconds_encode_plus_emit_overlapped": dt_both, "encode_seconds_while_emitting": dt_enc_overlapped,
            "docs_per_s_overlapped_whole_job": n_total / dt_both,
            "projected_seconds_for_1M_docs_with_bulk_bodies": 1e6 / (n_total / dt_both)}
    out["config5_bulk_encode"] = {"docs_total": n_total, "docs_per_gpu": n_local, "seq_len": 256, "top_k": 64,
                                  "emit": emit,
                                  "seconds": dt, "docs_per_s": n_total / dt, "mean_terms_emitted": nnz,
                                  "projected_seconds_for_1M_docs": 1e6 / (n_total / dt),
                                  "d2h_bytes_per_doc": 64 * 8 + 4,
                                  "trunk": "native packed-token trunk (splade_trunk_encode)" if args.native_trunk
                                  else "transformers trunk (reference) + fused head"}

    # ---------------- N1 check: what the un-fused head prologue (dense 768x768 -> GELU -> LayerNorm) costs
    x = torch.randn(64, 256, H, generator=g).to(dev)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ms_pro = cuda_time(lambda: model.model.head(x).to(torch.bfloat16), 20)
    out["head_prologue_hf"] = {"ms": ms_pro, "shape": "[64,256,768]",
                               "note": "HF ModernBertPredictionHead under autocast + the bf16 cast; fusing it into "
                                       "the projection (SURVEY 8(f) N1) could save at most this per encoder call"}

    # ---------------- config 3 again, full model (trunk + head + top-k)
    q_ids = torch.randint(6, V - 1, (Bq, Sq), generator=g).to(dev)
    q_mask = torch.ones(Bq, Sq, dtype=torch.int64, device=dev)

    def c3_full():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            rep = model.encode(q_ids, q_mask)
        return emitter.select(rep, k)

    ms = max_ranks(cuda_time(c3_full, 5))
    out["config3_full_model"] = {"batch": Bq, "seq_len": Sq, "k": k, "ms": ms, "queries_per_s": world * Bq / ms * 1e3}

    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
