"""Host-only timing of SparseEmitter.to_bulk (splade_emit_bulk): the `_bulk` request bodies of one 125 000-document
shard of BASELINE config 5 from synthetic compact top-64 buffers, sparse_vector format, 4-decimal weights - what
tools/bench_configs.py runs on a side thread behind the GPU encode.  Needs no GPU.
    python tools/time_emit.py [docs] [threads ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import splade_b200  # noqa: E402
from splade_b200._lib import load_library  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
threads = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8, 16]
k, chunk = 64, 8192
rng = np.random.default_rng(0)
ids = torch.from_numpy(rng.integers(0, 50000, size=(n_docs, k)).astype(np.int32))
vals = torch.from_numpy(np.sort(rng.gamma(2.0, 0.4, size=(n_docs, k)).astype(np.float32))[:, ::-1].copy())
counts = torch.full((n_docs,), k, dtype=torch.int32)
doc_ids = [str(i) for i in range(n_docs)]
em = splade_b200.SparseEmitter()
lib = load_library()


def emit_all():
    total = 0
    for c0 in range(0, n_docs, chunk):
        body = em.to_bulk(ids[c0:c0 + chunk], vals[c0:c0 + chunk], counts[c0:c0 + chunk], "bench-index",
                          doc_ids[c0:c0 + chunk], round_decimals=4, skip_empty=True)
        total += len(body)
    return total


for t in threads:
    prev = lib.splade_emit_set_threads(t)
    emit_all()
    t0 = time.perf_counter()
    nbytes = emit_all()
    dt = time.perf_counter() - t0
    lib.splade_emit_set_threads(prev)
    print(json.dumps({"docs": n_docs, "k": k, "emit_threads": t, "host_cores": os.cpu_count(), "seconds": round(dt, 3),
                      "docs_per_s": round(n_docs / dt), "MB_per_s": round(nbytes / dt / 1e6, 1),
                      "body_MB": round(nbytes / 1e6, 1)}), flush=True)
