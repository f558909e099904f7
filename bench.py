#!/usr/bin/env python
"""Benchmark of the SPLADE-max head hot path (BASELINE.json metric: SPLADE encode docs/sec at L=256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: BASELINE config 2, the fused
vocabulary projection + log1p(relu) + mask + max-pool of 64 documents x 256 tokens (hidden 768 -> 50 000
terms, 1.258 TFLOP).  With N GPUs every rank encodes its own 64-document shard (documents are independent:
no data-path collective, weak scaling); `value` is the whole-job docs/s = N*64*K / max-over-ranks time.

  value     device-resident: hidden states already in HBM when the timed region starts
  e2e       the same metric through the public API with HOST buffers: every step copies its hidden states
            and mask from pinned host memory, runs the fused head, and copies the [64, 50 000] fp32 sparse
            vectors back to pinned host memory (copies overlapped with compute on separate streams)
  roofline  the tcgen05 head kernel against the measured cuBLAS bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline / --impl reference
            the reference's CPU torch path (oracle/torch_reference.py, fp32, all host threads) on a bounded
            sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

# the bench falls back to NCCL when a peer-exchange barrier is not answered; do not wait NCCL-long for that
os.environ.setdefault("SPLADE_B200_PEER_TIMEOUT_MS", "20000")

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_DOCS, SEQ, HIDDEN, VOCAB = 64, 256, 768, 50000
WORKLOAD = "C2: SPLADE-max head encode, A.X-Encoder-base head shape (hidden 768 -> vocab 50000), batch 64 x L=256"
FLOP_PER_STEP = 2.0 * B_DOCS * SEQ * HIDDEN * VOCAB          # 76.8 MFLOP per token
N_ROTATE = 3                                                 # distinct input sets (see config.l2)
LAUNCHES_PER_STEP = 6                                        # kernels per splade_head_fwd call (see gpu_launches)


def bench_config(world: int) -> dict:
    """The `config` object of the JSON line - the SAME dict for our arm and for `--impl reference`."""
    return {"workload": WORKLOAD, "docs_per_gpu_per_step": B_DOCS, "seq_len": SEQ, "hidden": HIDDEN,
            "vocab": VOCAB, "sharding": f"documents sharded over {world} rank(s), no data-path collective",
            "l2": f"inputs larger than L2: rotating {N_ROTATE} distinct (hidden, W, bias) sets, "
                  f"{N_ROTATE} x 102 MB > 126 MB"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"burst": float(p["bf16_tflops"]), "sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "hbm_gbs": float(p["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the run through NVML (same counters as the nvidia-smi clocks line of the recipe)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x0000000000000004: "sw_power_cap",
        0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples = []  # (t, sm_mhz, reasons_mask)
        self.period = period_s
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:  # CUDA ordinal -> NVML handle through the UUID (robust to CUDA_VISIBLE_DEVICES remapping)
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # pragma: no cover
            self.nv, self.h, self.err = None, None, repr(e)

    def _run(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), float(mhz), int(mask)))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self.stop_flag.set()
        if self.thread is not None:
            self.thread.join(timeout=2)

    def summary(self, windows):
        """`windows` = [(t0, t1), ...] of time.perf_counter() stamps: the timed regions."""
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": "NVML unavailable"}
        win = [s for s in self.samples if any(t0 <= s[0] <= t1 for t0, t1 in windows)]
        note = "sampled inside the timed regions"
        if len(win) < 3:
            win = self.samples
            note = "timed regions shorter than the sampling period: all samples of the loaded run"
        mhz = sorted(s[1] for s in win)
        mask = 0
        for s in win:
            mask |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if mask & bit]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(win), "note": note}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU torch path on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_step_factory(docs_per_step: int, seed: int = 0):
    import torch

    from oracle import torch_reference as TR
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(docs_per_step, SEQ, HIDDEN, generator=g)
    W = torch.randn(VOCAB, HIDDEN, generator=g) * 0.0176
    bias = torch.randn(VOCAB, generator=g) * 0.02
    mask = torch.ones(docs_per_step, SEQ, dtype=torch.int64)

    def step():
        with torch.no_grad():
            return TR.head_chain(hidden, W, bias, mask)[0]

    return step


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def time_cpu_reference(docs_per_step: int, steps: int, warmup: int):
    import torch
    cores = host_cores()
    torch.set_num_threads(cores)
    step = cpu_reference_step_factory(docs_per_step)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return docs_per_step * steps / dt, dt / steps, cores


def time_cpu_full_encode(docs: int = 8, passes: int = 2):
    """the reference's WHOLE encode (22-layer A.X-Encoder-base-shaped trunk + head chain, fp32,
    no autocast - benchmark/encoders.py:309-318) on the host cores, on a bounded sample; the CPU counterpart of our
    arm's `e2e` (which also runs the whole model).  -> dict, or {"error": ...}."""
    try:
        import torch

        from oracle import torch_reference as TR
        cores = host_cores()
        torch.set_num_threads(cores)
        model = build_axbase_model()          # stays on the CPU; the drop-in class only wraps the HF model here
        ref = TR.ReferenceSpladeModule(model.model).eval()
        g = torch.Generator().manual_seed(1)
        ids = torch.randint(6, VOCAB - 1, (docs, SEQ), generator=g)
        mask = torch.ones(docs, SEQ, dtype=torch.int64)
        with torch.no_grad():
            ref(ids, mask)
            t0 = time.perf_counter()
            for _ in range(passes):
                ref(ids, mask)
            dt = (time.perf_counter() - t0) / passes
        return {"value": docs / dt, "unit": "docs/s", "cores": cores, "seconds_per_pass": dt,
                "sample": f"{docs} docs x L={SEQ} per pass, {passes} timed passes after 1 warm-up, fp32 torch CPU, the "
                          f"reference module chain (trunk + decoder + log1p/relu/mask/max)"}
    except Exception as e:  # noqa: BLE001 - context only
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    """`--impl reference`: rank 0 times the reference's own CPU implementation of the path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # size the per-step sample so that the whole run stays within ~2 minutes
    probe_docs = 4
    rate, t_step, cores = time_cpu_reference(probe_docs, steps=1, warmup=1)
    budget_s = 120.0
    per_step_budget = budget_s / max(1, args.steps + args.warmup)
    docs = int(max(1, min(B_DOCS, per_step_budget * rate)))
    value, t_step, cores = time_cpu_reference(docs, args.steps, args.warmup)
    sample = f"{docs} docs x L={SEQ} per step (of the 64-doc batch), fp32 torch CPU head chain, {cores} threads"
    line = {
        "impl": "reference", "metric": "splade_head_encode_docs_per_sec", "value": value, "unit": "docs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(int(os.environ.get("WORLD_SIZE", "1"))),
        "cpu_baseline": {"value": value, "unit": "docs/s", "cores": cores, "kind": "port", "sample": sample,
                         "full_encode": None if args.no_full_model else time_cpu_full_encode()},
        "e2e": {"value": value, "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------------
# secondary metric: training-step tokens/s of the hot path (BASELINE config 4)
# ------------------------------------------------------------------------------------------------------
def run_train_leg(dev, world, rank, steps, warmup, k_neg=1, gather=False, bias_offset=-1.0, exchange="peer",
                  grad_bf16=False, accum=1):
    """One C4 micro-step of the HOT PATH per rank: fused head forward for query (64x64), positive (64x256)
    and k negatives (64k x 256) hidden states, SPLADELossV33 (InfoNCE + FLOPS, optional cross-GPU negatives),
    backward through the loss and the sparse head backward, then the gradient all-reduce DDP would do for the
    tensors this path owns (tied decoder/embedding matrix 153.6 MB fp32 + bias).  The ModernBERT trunk is not
    part of the path (it stays the reference's) and is not timed here."""
    import torch
    import torch.distributed as dist

    import splade_b200
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    Bq = B_DOCS
    hq = torch.randn(Bq, 64, HIDDEN, generator=g).to(dev).requires_grad_(True)
    hp = torch.randn(Bq, SEQ, HIDDEN, generator=g).to(dev).requires_grad_(True)
    hn = torch.randn(Bq * k_neg, SEQ, HIDDEN, generator=g).to(dev).requires_grad_(True)
    W = (torch.randn(VOCAB, HIDDEN, generator=g) * 0.0176).to(dev).requires_grad_(True)       # fp32 master copy
    bias = (torch.randn(VOCAB, generator=g) * 0.02 + bias_offset).to(dev).requires_grad_(True)
    ones = lambda n, s: torch.ones(n, s, dtype=torch.int64, device=dev)  # noqa: E731
    mq, mp, mn = ones(Bq, 64), ones(Bq, SEQ), ones(Bq * k_neg, SEQ)
    # cross-GPU negatives: through NVLink peer memory inside the loss kernels (PeerExchange) unless asked for / forced
    # back to the NCCL all-gather + reduce-scatter statement; every rank takes the same branch
    exchange_used, exchange_obj = "none", None
    if gather and world > 1:
        exchange_used = "nccl"
        if exchange == "peer":
            try:  # construction fails on EVERY rank or on none (peer_exchange.py)
                exchange_obj = splade_b200.PeerExchange(Bq, VOCAB)
                exchange_used = "peer"
            except RuntimeError as e:  # e.g. CUDA IPC not permitted in this container
                exchange_obj = None
                exchange_used = f"nccl (peer exchange unavailable: {str(e)[:200]})"
    loss_box = [splade_b200.SPLADELossV33(gather_negatives=gather, exchange=exchange_obj)]

    def exchange_ok():
        """Collective: True when no rank saw a peer barrier time out (every rank gets the same answer)."""
        ok = 1.0
        try:
            exchange_obj.check()
        except RuntimeError:
            ok = 0.0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item() > 0.5)

    def step(i):
        loss_fn = loss_box[0]
        last = (i + 1) % accum == 0
        for t in (hq, hp, hn) + ((W, bias) if i % accum == 0 else ()):
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a, _ = splade_b200.splade_head(hq, W, bias, mq)
            p, _ = splade_b200.splade_head(hp, W, bias, mp)
            n, _ = splade_b200.splade_head(hn, W, bias, mn)
            if k_neg > 1:
                n = n.view(Bq, k_neg, -1)
            loss, d = loss_fn(a, p, n, global_step=1000 + i)
        (loss / accum if accum > 1 else loss).backward()
        if not last:
            return d
        if world > 1:
            if grad_bf16:  # N3: the tied matrix's gradient travels as bf16 (what DDP's bf16_compress_hook does)
                g16 = W.grad.to(torch.bfloat16)
                dist.all_reduce(g16, op=dist.ReduceOp.AVG)
                W.grad.copy_(g16)
            else:
                dist.all_reduce(W.grad, op=dist.ReduceOp.AVG)
            dist.all_reduce(bias.grad, op=dist.ReduceOp.AVG)
        with torch.no_grad():  # stand-in for the optimizer: bumps the version so the bf16 weight copy is re-cast
            W.add_(W.grad, alpha=-1e-6)
        return d

    warmup = -(-warmup // accum) * accum  # whole accumulation windows
    steps = -(-steps // accum) * accum
    for i in range(warmup):
        d = step(i)
    if exchange_obj is not None and not exchange_ok():
        # every rank switches to the NCCL statement of the exchange together and warms up again
        exchange_obj.close()
        exchange_obj = None
        exchange_used = "nccl (a peer-exchange barrier timed out during warm-up)"
        loss_box[0] = splade_b200.SPLADELossV33(gather_negatives=gather)
        for i in range(warmup):
            d = step(i)
    import gc
    gc.collect()
    gc.disable()   # no collector pause inside the timed region (a step is ~3 ms)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    # barrier + synchronize LAST: the collection above takes a different time on every rank, and a rank that enters
    # the timed region early spends the difference (measured: ~5 ms at N = 2) waiting in its first collective
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    marks[0].record()
    try:
        for i in range(steps):
            d = step(i)
            marks[i + 1].record()
    finally:
        gc.enable()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    exchange_failed = False
    if exchange_obj is not None:
        exchange_failed = not exchange_ok()
        exchange_obj.close()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if exchange_failed:
        raise RuntimeError("a peer-exchange barrier timed out inside the timed region; the training leg is invalid")
    per_step = sorted(round(marks[j].elapsed_time(marks[j + 1]), 3) for j in range(steps))
    # robust companion of the mean: per accumulation window (so the all-reduce micro-step is included), the median
    # window of every rank, max over ranks
    windows = sorted(marks[j].elapsed_time(marks[j + accum]) / accum for j in range(0, steps, accum))
    med = windows[len(windows) // 2]
    if world > 1:
        t = torch.tensor([med], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        med = float(t.item())
    tokens = world * Bq * (64 + SEQ + SEQ * k_neg) * steps
    head_flop = 2.0 * HIDDEN * VOCAB * Bq * (64 + SEQ + SEQ * k_neg)
    return {"tokens_per_s": tokens / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps,
            "ms_step_rank0_min_median_max": [per_step[0], per_step[len(per_step) // 2], per_step[-1]],
            "ms_per_step_median_window_max_over_ranks": med,
            "tokens_per_s_median_window": world * Bq * (64 + SEQ + SEQ * k_neg) / (med * 1e-3),
            "workload": f"C4 hot path per GPU: head fwd+bwd for q 64x64 / pos 64x256 / neg {64 * k_neg}x256 + "
                        f"SPLADELossV33 fwd+bwd (k={k_neg}, gather_negatives={gather}) + all-reduce of the tied "
                        f"decoder matrix ({'bf16-compressed' if grad_bf16 else 'fp32'}) and bias grads"
                        + (f" once per {accum} micro-steps (gradient accumulation, no_sync on the others)" if accum > 1
                           else "") + "; trunk excluded",
            "accum": accum,
            "head_fwd_tflop_per_step_per_gpu": head_flop / 1e12, "nonzero_d": d["nonzero_d"],
            # north-star figure: the WHOLE hot-path step (3 head forwards + sparse backward + loss forward / backward +
            # the gradient exchange) counted with the forward's algorithmic flop only, against the measured bf16 peak
            "head_fwd_tflops_over_whole_step": head_flop / 1e12 / (ms / steps * 1e-3),
            "fraction_of_bf16_peak_whole_step": head_flop / 1e12 / (ms / steps * 1e-3) / load_peaks()["burst"],
            "regime": f"decoder.bias offset {bias_offset}: {d['nonzero_d']:.0f} active terms per document",
            "negatives_exchange": exchange_used, "scaling": "weak"}


# ------------------------------------------------------------------------------------------------------
# context: full encode (unchanged ModernBERT trunk + head), ours vs the reference statement on the same GPU
# ------------------------------------------------------------------------------------------------------
def run_full_encode_leg(dev, steps=5, warmup=2, ours=None):
    """SPLADEModernBERT.encode on a random-init A.X-Encoder-base-shaped model (22 layers, hidden 768, vocab
    50 000; config restated from huggingface/v33/config.json), batch 64 x L=256, under autocast(bf16):
      ours               the reference's transformers trunk + the fused head (the drop-in default),
      ours_native_trunk  the native packed-token trunk + fused head in one C-ABI call (splade_trunk_encode),
      reference_chain    the reference's statement (HF logits -> relu -> log1p -> mask -> max)
    on the same device, plus how closely the native trunk's vectors agree with the transformers trunk's."""
    import torch

    import splade_b200
    from oracle import torch_reference as TR
    if ours is None:
        ours = build_axbase_model().to(dev).eval()
    ref = TR.ReferenceSpladeModule(ours.model).eval()
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(6, VOCAB - 1, (B_DOCS, SEQ), generator=g).to(dev)
    mask = torch.ones(B_DOCS, SEQ, dtype=torch.int64, device=dev)
    out = {"params": sum(p.numel() for p in ours.parameters()), "batch": B_DOCS, "seq_len": SEQ}
    native = splade_b200.NativeTrunk(ours.model)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        a, b = native.encode(ids, mask), ours.encode(ids, mask)
        ta, tb = a.topk(64, dim=1).indices.tolist(), b.topk(64, dim=1).indices.tolist()
        out["native_vs_transformers_trunk"] = {
            "cosine_min": float(torch.nn.functional.cosine_similarity(a, b, dim=1).min()),
            "max_abs_diff": float((a - b).abs().max()),
            "top64_overlap": sum(len(set(x) & set(y)) for x, y in zip(ta, tb)) / (64.0 * B_DOCS)}
        del a, b
        for name, fn in (("ours", lambda: ours.encode(ids, mask)),
                         ("ours_native_trunk", lambda: native.encode(ids, mask)),
                         ("reference_chain", lambda: ref(ids, mask)[0])):
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_batch": ms, "docs_per_s": B_DOCS / ms * 1e3}
    out["speedup"] = out["reference_chain"]["ms_per_batch"] / out["ours"]["ms_per_batch"]
    out["speedup_native_trunk"] = out["reference_chain"]["ms_per_batch"] / out["ours_native_trunk"]["ms_per_batch"]
    return out

def build_axbase_model(bias_offset=None, seed=0):
    """Random-init A.X-Encoder-base-shaped SPLADEModernBERT (22 layers, hidden 768, 12 heads, intermediate 1152,
    vocab 50 000, 149 M parameters; config restated from huggingface/v33/config.json) through the drop-in class's
    own constructor.  ``bias_offset`` shifts decoder.bias so that the pooled vectors are as sparse as a trained
    model's."""
    import tempfile

    import torch
    from transformers import AutoModelForMaskedLM, ModernBertConfig

    import splade_b200
    cfg = ModernBertConfig(vocab_size=VOCAB, hidden_size=HIDDEN, intermediate_size=1152, num_hidden_layers=22,
                           num_attention_heads=12, pad_token_id=VOCAB - 1, bos_token_id=0, eos_token_id=1,
                           cls_token_id=0, sep_token_id=1, max_position_embeddings=16384,
                           global_attn_every_n_layers=3, local_attention=128)
    torch.manual_seed(seed)
    hf = AutoModelForMaskedLM.from_config(cfg)
    if bias_offset is not None:
        with torch.no_grad():
            hf.decoder.bias.copy_(torch.randn(VOCAB, generator=torch.Generator().manual_seed(2)) * 0.02 + bias_offset)
    with tempfile.TemporaryDirectory() as d:
        hf.save_pretrained(d)
        return splade_b200.SPLADEModernBERT(model_name=d)


def run_e2e_encode_leg(dev, model, steps, warmup, world, barrier, max_over_ranks, native_trunk, top_k=64):
    """The reference's production encode (benchmark/encoders.py:309-345) end to end through the public API: token
    ids + attention mask in PINNED HOST memory -> H2D -> SPLADEModernBERT.encode under autocast(bf16) -> GPU top-k
    (splade_topk) -> ids / weights / counts back to PINNED HOST memory.  Every step's copies are inside the timed
    region; two slots on two streams so the copies of one batch overlap the compute of the other."""
    import torch

    import splade_b200
    g = torch.Generator().manual_seed(7)
    ids_h = [torch.randint(6, VOCAB - 1, (B_DOCS, SEQ), generator=g).pin_memory() for _ in range(2)]
    mask_h = torch.ones(B_DOCS, SEQ, dtype=torch.int64).pin_memory()
    out_h = [(torch.empty(B_DOCS, top_k, dtype=torch.int32).pin_memory(),
              torch.empty(B_DOCS, top_k, dtype=torch.float32).pin_memory(),
              torch.empty(B_DOCS, dtype=torch.int32).pin_memory()) for _ in range(2)]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    previous = getattr(model, "_native_trunk_enabled", False)
    model.use_native_trunk(native_trunk)

    def loop(n):
        for i in range(n):
            j = i & 1
            with torch.cuda.stream(streams[j]), torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                ids = ids_h[j].to(dev, non_blocking=True)
                mask = mask_h.to(dev, non_blocking=True)
                rep = model.encode(ids, mask)
                ti, tv, tc = splade_b200.topk(rep, top_k)
                out_h[j][0].copy_(ti, non_blocking=True)
                out_h[j][1].copy_(tv, non_blocking=True)
                out_h[j][2].copy_(tc, non_blocking=True)

    try:
        for st in streams:
            st.wait_stream(torch.cuda.current_stream(dev))
        loop(max(2, warmup))
        barrier()
        t0 = time.perf_counter()
        loop(steps)
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        barrier()
    finally:
        model.use_native_trunk(previous)
    h2d = ids_h[0].numel() * 8 + mask_h.numel() * 8
    d2h = sum(t.numel() * t.element_size() for t in out_h[0])
    j = (steps - 1) & 1
    return {"value": world * B_DOCS * steps / dt, "unit": "docs/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h, "steps": steps, "ms_per_step": dt / steps * 1e3,
            "trunk": "native packed-token trunk (splade_trunk_encode, opt-in)" if native_trunk
                     else "transformers ModernBERT trunk (the drop-in default)",
            "top_k": top_k, "checksum": float(out_h[j][1][0, :8].sum()), "counts0": int(out_h[j][2][0])}


def run_peer_parity(dev, world, rank, exchange="peer"):
    """Real-IPC parity of the NVLink peer exchange, run inside the N >= 2 bench so that the driver's SCALE record
    carries it: SPLADELossV33 with exchange=PeerExchange against the NCCL all-gather / reduce-scatter statement on
    the same seeded inputs (both window parities, autocast on and off, MarginMSE teachers included)."""
    import torch
    import torch.distributed as dist

    import splade_b200
    if world < 2 or exchange != "peer":
        return None
    try:
        x = splade_b200.PeerExchange(B_DOCS, VOCAB)
    except RuntimeError as e:
        return {"error": f"peer exchange unavailable: {str(e)[:200]}"}
    out = {"loss_bitexact": True, "dq_bitexact": True, "dnegs_bitexact": True, "ddocs_max_rel": 0.0,
           "loss_dict_equal": True}
    try:
        g = torch.Generator().manual_seed(500 + rank)

        def reps(n):
            return (torch.rand(n, VOCAB, generator=g) * (torch.rand(n, VOCAB, generator=g) < 0.02)).to(dev)

        q, p, n = reps(B_DOCS), reps(B_DOCS), reps(B_DOCS)
        t_pos = torch.randn(B_DOCS, generator=g).to(dev)
        t_neg = torch.randn(B_DOCS, generator=g).to(dev)
        f_peer = splade_b200.SPLADELossV33(exchange=x, lambda_margin_mse=0.5)
        f_nccl = splade_b200.SPLADELossV33(gather_negatives=True, lambda_margin_mse=0.5)

        def run(fn, autocast):
            a, b, c = (t.clone().requires_grad_(True) for t in (q, p, n))
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                loss, d = fn(a, b, c, global_step=20000, teacher_pos_scores=t_pos, teacher_neg_scores=t_neg)
            (loss / (1 + rank)).backward()
            return loss.detach(), d, a.grad, b.grad, c.grad

        for autocast in (True, False):
            for _ in range(2):  # both window parities
                got, want = run(f_peer, autocast), run(f_nccl, autocast)
                out["loss_bitexact"] &= bool(torch.equal(got[0], want[0]))
                out["dq_bitexact"] &= bool(torch.equal(got[2], want[2]))
                out["dnegs_bitexact"] &= bool(torch.equal(got[4], want[4]))
                rel = float((got[3] - want[3]).abs().max()) / (float(want[3].abs().max()) + 1e-30)
                out["ddocs_max_rel"] = max(out["ddocs_max_rel"], rel)
                out["loss_dict_equal"] &= all(got[1][k] == want[1][k] for k in want[1])
        x.check()
    except RuntimeError as e:
        out = {"error": f"{type(e).__name__}: {e}"[:300]}
    # every rank reports; rank 0's line carries the worst case
    flags = torch.tensor([float(out.get("loss_bitexact", False)), float(out.get("dq_bitexact", False)),
                          float(out.get("dnegs_bitexact", False)), float(out.get("loss_dict_equal", False)),
                          -float(out.get("ddocs_max_rel", 1.0))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    x.close()
    f = flags.tolist()
    res = {"loss_bitexact": f[0] > 0.5, "dq_bitexact": f[1] > 0.5, "dnegs_bitexact": f[2] > 0.5,
           "loss_dict_equal": f[3] > 0.5, "ddocs_max_rel": -f[4], "ranks": world,
           "what": "SPLADELossV33(exchange=PeerExchange) vs SPLADELossV33(gather_negatives=True) [NCCL all-gather + "
                   "reduce-scatter] on identical inputs, real CUDA-IPC windows; min / max over ranks; ddocs differs "
                   "by the summation order over ranks only"}
    if "error" in out:
        res["error"] = out["error"]
    return res


def run_train_full_leg(dev, world, rank, local_rank, accum=4, opt_steps=2, exchange="peer"):
    """The FULL training step of the reference (src/train/cli/train_v33_ddp.py:337-374) on the A.X-Encoder-base
    shape, BASELINE config 4: per GPU 64 queries x 64 + 64 positives x 256 + 64 negatives x 256 tokens through the
    22-layer transformers trunk WITH autograd, three head calls, SPLADELossV33, (loss / accum).backward() under
    DistributedDataParallel (:539-544) for N > 1, and - every `accum` micro-steps - clip_grad_norm_(1.0) + AdamW.
    decoder.bias is offset so the pooled vectors are sparse like a trained model's.  Variants:
      reference_semantics  drop-in classes, DDP all-reduces on EVERY micro-step (what the reference does)
      no_sync              DDP.no_sync() on the non-final micro-steps of the accumulation window (N3)
      bf16_grad_hook       + bf16 gradient compression of the all-reduce (N3)
      peer_negatives       + cross-GPU in-batch negatives through the NVLink peer exchange (N > 1)
      reference_statement  the reference's own head/loss statement (oracle/torch_reference.py) on the same trunk
      native_trunk         opt-in: the trunk too runs natively (NativeTrunkTrain: tcgen05 forward kernels, hand-driven
                           backward - DESIGN 3.9); same drop-in classes, same DDP semantics as reference_semantics
      native_trunk_fused_calls  + one encoder call per micro-step for queries, positives and negatives (N3)
    tokens/s = whole job (all ranks, real padded tokens of the collator shape) / max-over-ranks CUDA-event time."""
    import gc

    import torch
    import torch.distributed as dist

    import splade_b200
    from oracle import torch_reference as TR
    from splade_b200.training import (StepConfig, enable_bf16_gradient_compression, synthetic_triplet_batch,
                                      train_micro_step)
    B = B_DOCS
    warm = 2 * accum          # two optimizer steps of warm-up: DDP rebuilds its buckets after the first synchronised backward
    n_micro = warm + accum * opt_steps
    batches = [synthetic_triplet_batch(B, 64, SEQ, 1, VOCAB, VOCAB - 1, seed=1000 * (42 + rank) + i, ragged=False)
               for i in range(n_micro)]
    tokens_per_micro = world * B * (64 + SEQ + SEQ)
    out = {"workload": "C4 full training step per GPU: q 64x64 / pos 64x256 / neg 64x256 tokens, A.X-Encoder-base "
                       "shape (22L, 149 M params, transformers trunk with autograd) + 3 head calls + SPLADELossV33 + "
                       f"clip + AdamW, gradient accumulation {accum}, autocast(bf16), DDP when n_gpus > 1",
           "n_gpus": world, "accum": accum, "micro_steps_timed": accum * opt_steps, "micro_steps_warmup": 2 * accum,
           "tokens_per_micro_step": tokens_per_micro, "scaling": "weak"}
    base = build_axbase_model(bias_offset=-2.0)
    sched = splade_b200.SPLADELossV33()

    def ref_loss(anchor_repr, positive_repr, negative_repr, global_step, **kw):
        lam = [sched._lambda_schedule(global_step, x) for x in (sched.lambda_q, sched.lambda_d, sched.lambda_neg)]
        loss = TR.v33_loss(anchor_repr, positive_repr, negative_repr, *lam)
        return loss, {"nonzero_d": float((positive_repr > 0).sum(1).float().mean())}

    def time_variant(module, loss_fn, no_sync, bf16_hook, fuse_calls=False):
        model = module.to(dev).train()
        if world > 1:
            from torch.nn.parallel import DistributedDataParallel as DDP
            model = DDP(model, device_ids=[local_rank], find_unused_parameters=False)
            if bf16_hook:
                enable_bf16_gradient_compression(model)
        opt = torch.optim.AdamW(model.parameters(), lr=5e-5, weight_decay=0.01)
        scfg = StepConfig(gradient_accumulation_steps=accum, skip_allreduce_on_accumulation=no_sync,
                          fuse_encoder_calls=fuse_calls)
        gs, d = 20000, {}
        torch.cuda.reset_peak_memory_stats(dev)
        for i in range(warm):
            loss, d, gs = train_micro_step(model, loss_fn, opt, None, batches[i], i, gs, scfg, dev)
        # the eager 22-layer trunk is host-launch-bound: a generation-2 garbage collection in the middle of a
        # micro-step shows up as a 50-90 ms hiccup (seen at N = 2) - collect now, keep the collector off while timing
        gc.collect()
        gc.disable()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(n_micro - warm + 1)]
        if world > 1:  # after the collection, whose duration differs between ranks
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        marks[0].record()
        try:
            for i in range(warm, n_micro):
                loss, d, gs = train_micro_step(model, loss_fn, opt, None, batches[i], i, gs, scfg, dev)
                marks[i - warm + 1].record()
        finally:
            gc.enable()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (n_micro - warm)
        per_step = [round(marks[j].elapsed_time(marks[j + 1]), 2) for j in range(n_micro - warm)]
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        res = {"ms_per_micro_step": ms, "tokens_per_s": tokens_per_micro / ms * 1e3,
               "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30, "last_loss": float(loss.detach()),
               "nonzero_d": d.get("nonzero_d"), "ms_each_micro_step_rank0": per_step,
               "median_ms_rank0": sorted(per_step)[len(per_step) // 2]}
        del opt, model
        module.zero_grad(set_to_none=True)
        torch.cuda.empty_cache()
        return res

    variants = [("reference_semantics", "ours", False, False, "local")]
    if world > 1:
        variants += [("no_sync", "ours", True, False, "local"), ("bf16_grad_hook", "ours", True, True, "local"),
                     ("peer_negatives", "ours", True, False, exchange)]
    variants.insert(1, ("native_trunk", "ours_native", False, False, "local"))
    # + ONE encoder call per micro-step for queries, positives and negatives (StepConfig.fuse_encoder_calls, N3)
    variants.insert(2, ("native_trunk_fused_calls", "ours_native", False, False, "local"))
    if world > 1:
        variants.insert(3, ("native_trunk_no_sync", "ours_native", True, False, "local"))
    variants.append(("reference_statement", "ref", False, False, "local"))
    ref_module = None
    for name, side, no_sync, hook, negs in variants:
        x = None
        try:
            if side in ("ours", "ours_native"):
                base.use_native_trunk(False, train=(side == "ours_native"))
                module = base
                loss_fn = splade_b200.SPLADELossV33()
                if negs == "peer":
                    x = splade_b200.PeerExchange(B, VOCAB)
                    loss_fn = splade_b200.SPLADELossV33(exchange=x)
                elif negs == "nccl":
                    loss_fn = splade_b200.SPLADELossV33(gather_negatives=True)
            else:
                if ref_module is None:
                    base.to("cpu")
                    torch.cuda.empty_cache()
                    ref_module = TR.ReferenceSpladeModule(base.model)
                module, loss_fn = ref_module, ref_loss
            out[name] = time_variant(module, loss_fn, no_sync, hook, fuse_calls=name.endswith("fused_calls"))
        except Exception as e:  # noqa: BLE001 - keep the other variants
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if x is not None:
            try:
                x.check()
                x.close()
            except RuntimeError as e:
                out[name]["exchange_error"] = str(e)[:200]
    # the collective that bounds the step: DDP's bucketed fp32 all-reduce of every gradient (597 MB), timed alone
    if world > 1:
        nbytes = sum(p.numel() for p in base.parameters()) * 4
        buf = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        tied = torch.empty(VOCAB * HIDDEN, dtype=torch.float32, device=dev)
        coll = {}
        for label, t in (("all_gradients_fp32", buf), ("tied_decoder_matrix_fp32", tied),
                         ("all_gradients_bf16", buf.to(torch.bfloat16))):
            for _ in range(2):
                dist.all_reduce(t)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                dist.all_reduce(t)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            sz = t.numel() * t.element_size()
            coll[label] = {"bytes": sz, "ms": ms, "busbw_GBps": sz * 2 * (world - 1) / world / ms / 1e6}
        out["collectives_timed_alone"] = coll
        a, b = out.get("reference_semantics", {}), out.get("no_sync", {})
        if "ms_each_micro_step_rank0" in a and "ms_each_micro_step_rank0" in b:
            # micro-steps that do not end an accumulation window: with no_sync they issue no collective at all
            inner = [j for j in range(accum * opt_steps) if (j + 1) % accum != 0]
            mean = lambda xs, idx: sum(xs[j] for j in idx) / len(idx)  # noqa: E731
            exposed = mean(a["ms_each_micro_step_rank0"], inner) - mean(b["ms_each_micro_step_rank0"], inner)
            out["limiting_collective"] = {
                "name": f"DDP's bucketed all-reduce of every gradient: {nbytes / 1e6:.0f} MB fp32 per micro-step, NCCL "
                        f"ring/NVLS over NVLink, {world} ranks",
                "ms_timed_alone": coll["all_gradients_fp32"]["ms"],
                "exposed_ms_per_micro_step": exposed,
                "how": "mean over the micro-steps that do not end an accumulation window of (reference_semantics - "
                       "no_sync), rank 0 CUDA events: with no_sync those micro-steps issue no collective",
                "note": "the tied decoder / embedding matrix is the LAST gradient of a backward pass (its last producer "
                        f"is the embedding lookup), so its bucket's all-reduce ({coll['tied_decoder_matrix_fp32']['ms']:.2f} "
                        "ms alone) can never overlap the backward; everything before it overlaps the 22-layer trunk"}
    a, r = out.get("reference_semantics", {}), out.get("reference_statement", {})
    if "ms_per_micro_step" in a and "ms_per_micro_step" in r:
        out["speedup_vs_reference_statement"] = r["ms_per_micro_step"] / a["ms_per_micro_step"]
    nt = out.get("native_trunk", {})
    if "ms_per_micro_step" in nt and "ms_per_micro_step" in r:
        out["speedup_native_trunk_vs_reference_statement"] = r["ms_per_micro_step"] / nt["ms_per_micro_step"]
    base.use_native_trunk(False, train=False)
    del base, ref_module
    torch.cuda.empty_cache()
    return out


def run_ragged_leg(dev, iters=20):
    """Right-padded batches as the reference's collator produces them (dynamic padding to the longest member,
    src/train/data/dataloader.py:46-164): the padded tiling vs the packed variable-length tiling
    (splade_head_fwd_mode) vs the device-side auto plan, for documents (C2 shape) and queries (C3 shape).
    FLOP are counted on REAL tokens only."""
    import torch

    import splade_b200
    out = {}
    g = torch.Generator().manual_seed(3)
    W = (torch.randn(VOCAB, HIDDEN, generator=g) * 0.0176).to(torch.bfloat16).to(dev)
    bias = (torch.randn(VOCAB, generator=g) * 0.02).to(torch.bfloat16).float().to(dev)
    cases = (("docs_64x256_len_U[32,256]", 64, 256, 32, 256), ("queries_512x64_len_U[4,24]", 512, 64, 4, 24))
    for name, B, S, lo, hi in cases:
        hidden = torch.randn(B, S, HIDDEN, generator=g).to(torch.bfloat16).to(dev)
        lens = torch.randint(lo, hi + 1, (B,), generator=g)
        mask = (torch.arange(S)[None, :] < lens[:, None]).to(torch.int32).to(dev)
        toks = int(lens.sum())
        row = {"real_tokens": toks, "real_frac": toks / (B * S)}
        plan = torch.zeros(4, dtype=torch.int32, device=dev)
        for label, mode in (("padded", 0), ("packed", 1), ("auto", -1)):
            def fn():
                return splade_b200.head_forward(hidden, W, bias, mask, False, pack_mode=mode, plan_out=plan)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            row[label + "_ms"] = ms
            row[label + "_tflops_on_real_tokens"] = 2.0 * toks * HIDDEN * VOCAB / ms / 1e9
        row["auto_picked"] = "packed" if int(plan[0]) == 1 else "padded"
        row["docs_per_s_auto"] = B / row["auto_ms"] * 1e3
        row["speedup_packed_over_padded"] = row["padded_ms"] / row["packed_ms"]
        out[name] = row
        del hidden
    return out


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import splade_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    numa_bound = splade_b200.distributed.bind_to_device_numa_node(local_rank) if world > 1 else False
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner out of stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    K, Wm = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- synthetic inputs: hidden ~ N(0,1), W ~ N(0, 0.0176^2), bias ~ N(0, 0.02), full lengths
    g = torch.Generator(device="cpu").manual_seed(42 + rank)
    sets = []
    for i in range(N_ROTATE):
        hidden = torch.randn(B_DOCS, SEQ, HIDDEN, generator=g).to(torch.bfloat16).to(dev)
        W = (torch.randn(VOCAB, HIDDEN, generator=g) * 0.0176).to(torch.bfloat16).to(dev)
        bias = (torch.randn(VOCAB, generator=g) * 0.02).to(torch.bfloat16).float().to(dev)
        mask = torch.ones(B_DOCS, SEQ, dtype=torch.int32, device=dev)
        sets.append((hidden, W, bias, mask))

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- (1) device-resident throughput + per-launch kernel time
    def step(i):
        h, w, b, m = sets[i % N_ROTATE]
        return splade_b200.head_forward(h, w, b, m, want_token_weights=False)

    for i in range(Wm):
        step(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    t_load0 = time.perf_counter()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(K):
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    end.record()
    barrier()
    t_load1 = time.perf_counter()
    total_ms = max_over_ranks(start.elapsed_time(end))
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / K
    value = world * B_DOCS * K / (total_ms * 1e-3)

    # ---- (2) end to end with host buffers: ONE C-ABI call per step (splade_head_fwd_host) enqueues
    #      H2D(hidden, mask) -> fused head -> D2H(sparse_repr); two slots on two streams so that the copies of one
    #      step overlap the compute of the other.  Every step's copies are inside the timed region.
    n_e2e = K
    host_in = [torch.randn(B_DOCS, SEQ, HIDDEN, generator=g).to(torch.bfloat16).pin_memory() for _ in range(2)]
    host_mask = torch.ones(B_DOCS, SEQ, dtype=torch.int32).pin_memory()
    host_out = [torch.empty(B_DOCS, VOCAB, dtype=torch.float32).pin_memory() for _ in range(2)]
    slots = [splade_b200.HostEncodeSlot(B_DOCS, SEQ, HIDDEN, VOCAB, dev) for _ in range(2)]
    streams = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
    w16, b32 = sets[0][1], sets[0][2]

    def e2e_loop(n):
        for i in range(n):
            j = i & 1
            splade_b200.head_forward_host(host_in[j], host_mask, w16, b32, slots[j], host_out[j], None, streams[j])

    for st in streams:
        st.wait_stream(torch.cuda.current_stream(dev))
    e2e_loop(Wm)
    barrier()
    t_load2 = time.perf_counter()
    t0 = time.perf_counter()
    e2e_loop(n_e2e)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clock_windows = [(t_load0, t_load1), (t_load2, time.perf_counter())]
    e2e_value = world * B_DOCS * n_e2e / e2e_s
    h2d_bytes = host_in[0].numel() * 2 + host_mask.numel() * 4
    d2h_bytes = host_out[0].numel() * 4
    checksum = float(host_out[(n_e2e - 1) & 1][0, :8].sum())  # touch the result on the host

    e2e_head = {"value": e2e_value, "unit": "docs/s", "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "steps": n_e2e,
                "api": "splade_b200.head_forward_host = one C-ABI call splade_head_fwd_host per step on pinned "
                       "host buffers holding HIDDEN STATES (H2D, fused head, D2H of the dense [64, 50 000] fp32 "
                       "vectors; 2 slots / 2 streams) - round 1's e2e, PCIe-bound at 8 GPUs: hidden states never "
                       "live on the host in the reference",
                "checksum": checksum, "numa_bound": numa_bound}
    del host_in, host_out, slots

    # ---- (2b) the reference's real encode boundary (benchmark/encoders.py:309-345): token ids in, top-k out
    e2e, e2e_hf, model, e2e_err = None, None, None, None
    if not args.no_full_model:
        try:
            model = build_axbase_model().to(dev).eval()
            t_a = time.perf_counter()
            e2e = run_e2e_encode_leg(dev, model, K, Wm, world, barrier, max_over_ranks, native_trunk=True)
            e2e_hf = run_e2e_encode_leg(dev, model, max(3, min(K, 20)), 2, world, barrier, max_over_ranks,
                                        native_trunk=False)
            clock_windows.append((t_a, time.perf_counter()))
        except Exception as e:  # noqa: BLE001 - never lose the headline line; every rank fails alike (same code)
            e2e_err = f"{type(e).__name__}: {e}"[:300]
    # the clock / throttle samples cover the headline and the e2e regions (clock_windows); the NVML polling thread
    # (one per rank, every 10 ms) stops here so that it cannot perturb the host-bound training legs below
    sampler.stop()

    train = None
    if not args.no_train:
        del sets[1:]
        gather = world > 1 and not args.local_negatives
        n_train = max(3, min(20, K))
        try:
            train = run_train_leg(dev, world, rank, steps=n_train, warmup=3, k_neg=1, gather=gather,
                                  exchange=args.exchange)
            # the same step in the trained-like sparse regime (bias offset sweeps the density)
            train["sparse_regime"] = run_train_leg(dev, world, rank, steps=n_train, warmup=3, k_neg=1, gather=gather,
                                                   bias_offset=-2.0, exchange=args.exchange)
            # N3 variant: gradient accumulation over 4 micro-steps with the all-reduce on the last one only
            train["sparse_regime_accum4_no_sync"] = run_train_leg(
                dev, world, rank, steps=n_train, warmup=4, k_neg=1, gather=gather, bias_offset=-2.0,
                exchange=args.exchange, accum=4)
            if world > 1:  # N3 variant: the same sparse-regime step with the gradient all-reduce in bf16
                train["sparse_regime_bf16_allreduce"] = run_train_leg(
                    dev, world, rank, steps=n_train, warmup=3, k_neg=1, gather=gather, bias_offset=-2.0,
                    exchange=args.exchange, grad_bf16=True)
        except RuntimeError as e:  # secondary leg: never lose the headline line over it
            train = dict(train or {}, error=f"{type(e).__name__}: {e}"[:400])

    peer_parity = None
    if world > 1 and not args.no_train and not args.local_negatives:
        try:
            peer_parity = run_peer_parity(dev, world, rank, args.exchange)
        except Exception as e:  # noqa: BLE001
            peer_parity = {"error": f"{type(e).__name__}: {e}"[:300]}

    train_full = None
    if not args.no_train_full:
        if model is not None:
            model.to("cpu")
        torch.cuda.empty_cache()
        try:
            train_full = run_train_full_leg(dev, world, rank, local_rank, exchange=args.exchange)
        except Exception as e:  # noqa: BLE001
            train_full = {"error": f"{type(e).__name__}: {e}"[:300]}
        if model is not None and world == 1:
            model.to(dev)

    gpu_ref = None
    if world == 1 and rank == 0 and not args.no_gpu_reference:
        # the fair GPU baseline: the reference's own chain (cuBLAS bf16 GEMM + ATen relu/log1p/mul/max under
        # autocast, oracle/torch_reference.py) on the same device and the same inputs
        try:
            from oracle import torch_reference as TR
            h, w, b, m = sets[0]
            hf, wf = h.float(), w.float()
            with torch.no_grad():
                for _ in range(2):
                    TR.head_chain_autocast(hf, wf, b, m)
                torch.cuda.synchronize()
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record()
                for _ in range(5):
                    TR.head_chain_autocast(hf, wf, b, m)
                r1.record()
                torch.cuda.synchronize()
            ref_ms = r0.elapsed_time(r1) / 5
            gpu_ref = {"ms_per_step": ref_ms, "docs_per_s": B_DOCS / ref_ms * 1e3,
                       "what": "reference head chain under torch.autocast(bf16) on the same GPU (cuBLAS + ATen)",
                       "speedup_of_fused_head": ref_ms / (total_ms / K)}
            del hf, wf
            torch.cuda.empty_cache()
        except Exception as e:
            gpu_ref = {"error": repr(e)[:200]}

    ragged = None
    if world == 1 and rank == 0 and not args.no_ragged:
        try:
            ragged = run_ragged_leg(dev)
        except Exception as e:  # context only
            ragged = {"error": repr(e)[:200]}

    full = None
    if world == 1 and rank == 0 and not args.no_full_model:
        try:
            full = run_full_encode_leg(dev, ours=model)
        except Exception as e:  # context only: never fail the bench line over it
            full = {"error": repr(e)[:200]}

    clocks = sampler.summary(clock_windows)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- (3) roofline + cpu baseline (rank 0, N=1 leg of the contract)
    peaks = load_peaks()
    timed_s = total_ms * 1e-3
    peak_kind = "burst" if timed_s < 2.0 else "sustained"
    achieved = FLOP_PER_STEP / (kernel_ms * 1e-3) / 1e12
    traffic = None
    prof = os.path.join(ROOT, "profiles", "head_fwd_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peaks[peak_kind], "unit": "TFLOP/s",
                "frac": achieved / peaks[peak_kind], "traffic": traffic, "kernel": "head_fwd_kernel<false>",
                "kernel_ms": kernel_ms, "peak_kind": f"{peak_kind}, {peaks['source']}",
                "algorithmic_flop_per_launch": FLOP_PER_STEP}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        docs = 8
        v, t_step, cores = time_cpu_reference(docs, steps=3, warmup=1)
        if t_step * 4 < 3.0:  # fast host: use a larger sample to reach ~10-30 s of CPU work
            docs = min(B_DOCS, int(docs * 3.0 / max(t_step, 1e-3)))
            v, t_step, cores = time_cpu_reference(docs, steps=3, warmup=1)
        cpu = {"value": v, "unit": "docs/s", "cores": cores, "kind": "port",
               "sample": f"{docs} docs x L={SEQ} per pass, 3 timed passes after 1 warm-up, fp32 torch CPU head chain "
                         f"(oracle/torch_reference.py)",
               # the CPU counterpart of `e2e` (whole 22-layer model, not only the head)
               "full_encode": None if args.no_full_model else time_cpu_full_encode()}

    line = {
        "metric": "splade_head_encode_docs_per_sec", "value": value, "unit": "docs/s", "n_gpus": world,
        "steps": K, "warmup": Wm, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": bench_config(world),
        "tokens_per_s": value * SEQ,
        "clocks": clocks,
        # e2e = the reference's production encode boundary: token ids (pinned host) -> SPLADEModernBERT.encode ->
        # top-k ids / weights (pinned host); e2e_head = the same head call as `value` fed from host hidden states
        "e2e": dict(e2e, api="token ids + mask in pinned host memory -> SPLADEModernBERT.encode (22-layer A.X-base "
                             "shape, autocast bf16) -> splade_topk(k=64) -> ids / weights / counts in pinned host "
                             "memory; 2 slots / 2 streams; every copy inside the timed region")
               if e2e is not None else dict(e2e_head, note=f"full-encode e2e unavailable: {e2e_err or 'skipped'}"),
        "e2e_reference_trunk": e2e_hf,
        "e2e_head": e2e_head,
        # kernels of libsplade_b200.so launched per step in the device-resident timed region (splade_head_fwd, auto
        # tiling policy): build_tile_masks, varlen_plan (counts + plans), varlen_pack (exits: padded tiling chosen),
        # head_fwd_kernel, head_fwd_varlen_kernel (exits), finalize_varlen (exits).  The e2e region adds 2 per step
        # (build_tile_masks + head_fwd_kernel: splade_head_fwd_host decides the tiling on the host).
        "gpu_launches": LAUNCHES_PER_STEP * K,
        "gpu_launches_per_step": LAUNCHES_PER_STEP,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "train": train,
        "train_full": train_full,
        "peer_parity": peer_parity,
        "ragged": ragged,
        "full_encode": full,
        "gpu_reference_chain": gpu_ref,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the secondary training-step leg")
    ap.add_argument("--no-full-model", action="store_true", help="skip the trunk+head legs (e2e falls back to e2e_head)")
    ap.add_argument("--no-train-full", action="store_true", help="skip the full DDP training-step leg")
    ap.add_argument("--no-ragged", action="store_true", help="skip the ragged-batch (packed tiling) context leg")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the reference-chain-on-GPU context leg")
    ap.add_argument("--gather-negatives", action="store_true",
                    help="(default for N > 1, kept for old command lines) training leg: every rank's positives are "
                         "in-batch negatives for every query (BASELINE config 4, north-star extension)")
    ap.add_argument("--local-negatives", action="store_true",
                    help="training leg: per-rank negatives only, the reference's own semantics (no exchange step)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="how the positives travel for N > 1: 'peer' = NVLink peer-memory loads inside the loss "
                         "kernels (PeerExchange), 'nccl' = all-gather forward / reduce-scatter backward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
