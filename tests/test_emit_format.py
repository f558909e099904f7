"""CPU: the emitters' exact integer number formatting (csrc/emit.cu) prints the same bytes as printf("%.{d}f") /
printf("%.9g") - checked against Python's % operator, which is correctly rounded like glibc - and the threaded
document ranges produce the same bytes as one thread.  (The formats themselves are pinned against the reference's
bulk_actions_sparse / NeuralSparseSearcher in tests/test_reference_dropin_cpu.py.)"""
import ctypes

import numpy as np
import pytest


def _emit_json(ids, vals, counts, decimals):
    from splade_b200._lib import load_library
    lib = load_library()
    B, k = ids.shape
    offsets = np.zeros(B + 1, dtype=np.int64)
    need = ctypes.c_size_t(0)
    cap = B * (2 + k * 48)
    out = np.empty(cap, dtype=np.uint8)
    rc = lib.splade_emit_json(ids.ctypes.data, vals.ctypes.data, counts.ctypes.data, B, k, None, None, 0, decimals,
                              out.ctypes.data, cap, offsets.ctypes.data, ctypes.byref(need))
    assert rc == 0, rc
    raw = out[: need.value].tobytes()
    return [raw[offsets[b]:offsets[b + 1]] for b in range(B)]


def _special_values():
    f = np.float32
    v = [0.125, 0.375, 0.5, 1.5, 2.5, 0.00005, 0.00015, 0.99995, 0.99999994, 9.9999, 99999.9999, 1e-4, 9.9999e-5,
         1.0000001e-4, 1e9, 9.9999994e8, 999999.94, 123456.79, 1.0, 10.0, 100.0, 0.1, 0.2, 0.3, 3.4e38, 1e-38, 1e-45,
         1.17549435e-38, 16777216.0, 33554432.0, 4.5, 0.045, 0.0045, 0.00045, 0.000045, 7.0, 1e7, 1e8, 2.5e8,
         0.30000001192092896, 0.05, 0.15, 0.25, 0.35, 1.00005, 1.00015, 8.5, 1234.5678, 0.00010001, 5e-5, 4.9999e-5]
    v += [np.nextafter(f(x), f(np.inf)) for x in (0.00005, 0.5, 1.0, 1e-4, 0.99995)]
    v += [np.nextafter(f(x), f(0)) for x in (0.00005, 0.5, 1.0, 1e-4, 0.99995, 1e9)]
    return np.array(v, dtype=np.float32)


@pytest.mark.parametrize("decimals", [-1, 0, 2, 4, 9])
def test_number_formats_equal_printf(decimals):
    rng = np.random.default_rng(decimals + 7)
    rand = np.exp(rng.uniform(np.log(1e-7), np.log(5e4), size=20000)).astype(np.float32)
    vals = np.concatenate([_special_values(), rand])
    k = 64
    pad = (-len(vals)) % k
    vals = np.concatenate([vals, np.ones(pad, np.float32)]).reshape(-1, k)
    B = vals.shape[0]
    ids = np.tile(np.arange(k, dtype=np.int32), (B, 1))
    counts = np.full(B, k, dtype=np.int32)
    docs = _emit_json(ids, np.ascontiguousarray(vals), counts, decimals)
    n_checked = 0
    for b, doc in enumerate(docs):
        text = doc.decode()
        assert text[0] == "{" and text[-1] == "}"
        got = dict(item.split(": ") for item in text[1:-1].split(", ")) if len(text) > 2 else {}
        for j in range(k):
            x = float(vals[b, j])
            want = ("%.9g" % x) if decimals < 0 else ("%.*f" % (decimals, x))
            key = '"%d"' % j
            if decimals >= 0 and float(want) <= 0.0:
                assert key not in got, (x, want)     # the rounded format drops what rounds to zero
                continue
            assert got[key] == want, (x, decimals, got[key], want)
            n_checked += 1
    assert n_checked > 5000


def test_threaded_ranges_equal_one_thread():
    from splade_b200._lib import load_library
    lib = load_library()
    rng = np.random.default_rng(3)
    B, k = 5000, 32
    ids = rng.integers(0, 50000, size=(B, k), dtype=np.int32)
    vals = np.exp(rng.uniform(np.log(1e-5), np.log(10), size=(B, k))).astype(np.float32)
    counts = rng.integers(0, k + 1, size=B, dtype=np.int32)
    prev = lib.splade_emit_set_threads(1)
    try:
        one = _emit_json(ids, vals, counts, 4)
        lib.splade_emit_set_threads(7)
        many = _emit_json(ids, vals, counts, 4)
    finally:
        lib.splade_emit_set_threads(prev)
    assert one == many
    assert sum(len(d) for d in one) > B * 10


def test_to_bulk_threaded_equals_one_thread_and_parses():
    import json

    import torch

    import splade_b200
    from splade_b200._lib import load_library
    lib = load_library()
    rng = np.random.default_rng(5)
    B, k = 3000, 16
    ids = torch.from_numpy(rng.integers(0, 1000, size=(B, k)).astype(np.int32))
    vals = torch.from_numpy(np.sort(rng.uniform(0.0, 3.0, size=(B, k)).astype(np.float32))[:, ::-1].copy())
    counts = torch.from_numpy(rng.integers(0, k + 1, size=B).astype(np.int32))
    em = splade_b200.SparseEmitter()
    doc_ids = [f"d{i}" for i in range(B)]
    sources = [{"content": f"text \"{i}\""} for i in range(B)]
    prev = lib.splade_emit_set_threads(1)
    try:
        one = em.to_bulk(ids, vals, counts, "idx", doc_ids, sources=sources, round_decimals=4, skip_empty=True)
        lib.splade_emit_set_threads(5)
        many = em.to_bulk(ids, vals, counts, "idx", doc_ids, sources=sources, round_decimals=4, skip_empty=True)
    finally:
        lib.splade_emit_set_threads(prev)
    assert one == many
    lines = one.decode().splitlines()
    assert len(lines) % 2 == 0 and len(lines) > B
    for a, s in zip(lines[0:200:2], lines[1:200:2]):
        act, src = json.loads(a), json.loads(s)
        i = int(act["index"]["_id"][1:])
        n = int(counts[i])
        want = {str(int(t)): float("%.4f" % float(v)) for t, v in zip(ids[i, :n], vals[i, :n])
                if float("%.4f" % float(v)) > 0}
        assert src["sparse_embedding"] == pytest.approx(want) and src["content"] == f'text "{i}"'
