"""CPU: the C-ABI shared library is present, loads, and exports every symbol include/splade_b200.h declares.
No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "splade_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(splade_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for s in ("splade_head_fwd", "splade_head_fwd_host", "splade_head_bwd", "splade_flops_nnz", "splade_scores", "splade_rowdot",
              "splade_v33_scalars", "splade_v33_loss_fwd", "splade_loss_bwd", "splade_topk", "splade_workspace_bytes",
              "splade_last_error", "splade_abi_version"):
        assert s in syms


def test_library_loads_and_exports_every_declared_symbol():
    import splade_b200
    from splade_b200 import _lib
    assert os.path.exists(splade_b200.LIB_PATH), "run __graft_entry__.build() first"
    lib = splade_b200.load_library()
    declared = _declared_symbols()
    assert set(declared) == set(_lib.EXPORTS), "python binding and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.splade_abi_version() == splade_b200.ABI_VERSION
    # pure host helper: workspace sizing needs no device
    assert lib.splade_workspace_bytes(_lib.OP_HEAD_FWD, 64, 256, 768, 50000, 0) >= 64 * 8 * 4   # tile validity bitmaps
    assert lib.splade_workspace_bytes(_lib.OP_V33_LOSS_FWD, 64, 64, 0, 50000, 1) > 64 * 64 * 4
    assert lib.splade_workspace_bytes(_lib.OP_HEAD_FWD_TOKW, 64, 256, 768, 50000, 0) >= 64 * 256 * 4
    assert lib.splade_workspace_bytes(_lib.OP_HEAD_FWD, 2, 300, 64, 260, 0) >= 2 * 260 * 4
    assert lib.splade_workspace_bytes(_lib.OP_HEAD_BWD, 64, 256, 768, 50000, 0) >= 2 * 64 * 50000 * 4


def test_struct_layout_matches_header():
    import ctypes
    from splade_b200._lib import V33FwdArgs, V33ScalarArgs
    # 8 int32 + 8 float + 11 pointers, no padding on LP64
    assert ctypes.sizeof(V33ScalarArgs) == 8 * 4 + 8 * 4 + 11 * 8
    # 8 int32 + 8 float + 11 pointers + size_t
    assert ctypes.sizeof(V33FwdArgs) == 8 * 4 + 8 * 4 + 11 * 8 + 8


def test_peer_exchange_sizing_and_struct():
    """Host-only pieces of the NVLink peer exchange: struct mirror, window and workspace sizes (no device call)."""
    import ctypes
    import splade_b200
    from splade_b200 import _lib
    lib = splade_b200.load_library()
    assert ctypes.sizeof(_lib.PeerGroup) == 4 * 4 + _lib.MAX_PEERS * 8
    B, V, W = 64, 50000, 8
    rep, g = B * V * 4, B * W * B * 4
    # header + 2 parities x (Q, P, G, gl), every region 256-byte aligned
    assert lib.splade_peer_window_bytes(B, V, W) == 512 + 2 * (2 * rep + g + 256)
    assert lib.splade_peer_window_bytes(B, V, _lib.MAX_PEERS + 1) == 0
    # gloc [W,B,B] fp32 + the bf16 copies of G [B, W*B] and G^T [W*B, B] the tensor-core dq works on
    assert lib.splade_workspace_bytes(_lib.OP_PEER_LOSS_BWD, B, 0, 0, 0, W) == W * B * B * 4 + 2 * (B * W * B * 2)
    assert lib.splade_workspace_bytes(_lib.OP_LOSS_BWD_TC, B, 0, 0, 0, W * B) == 2 * (B * W * B * 2)
    with pytest.raises(RuntimeError, match="process group"):
        splade_b200.PeerExchange(B, V)


def test_product_refuses_cpu_tensors():
    import torch
    import splade_b200
    h = torch.zeros(1, 4, 64, dtype=torch.bfloat16)
    w = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        splade_b200.head_forward(h, w, torch.zeros(8), torch.ones(1, 4, dtype=torch.int32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        splade_b200.topk(torch.rand(2, 16), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        splade_b200.SPLADELossV33()(torch.rand(2, 8), torch.rand(2, 8), torch.rand(2, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "opensearch-neural-pre-train_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_packed_layout_and_trunk_sizes_without_a_gpu():
    """Host-only entry points of the packed layout / native trunk: row counts, workspace sizes, argument checks."""
    import ctypes

    import splade_b200
    from splade_b200 import _lib as L
    lib = splade_b200.load_library()
    # every document rounded up to 8 tokens, whole 256-row tiles
    assert lib.splade_packed_rows(64, 256) == 16384
    assert lib.splade_packed_rows(3, 7) == 256
    assert lib.splade_packed_rows(512, 60) == 512 * 64
    head = lib.splade_workspace_bytes(L.OP_HEAD_FWD, 64, 256, 768, 50000, 0)
    assert head >= 64 * 256 * 768 * 2 + 64 * 50000 * 4          # packed token copy + merge keys
    assert lib.splade_workspace_bytes(L.OP_HEAD_FWD_TOKW, 64, 256, 768, 50000, 0) >= head + 64 * 256 * 4
    trunk = lib.splade_workspace_bytes(L.OP_TRUNK_ENCODE, 64, 256, 768, 50000, 1152)
    rows = 16384
    assert trunk >= rows * (768 * 4 + 768 * 2 + 3 * 768 * 2 + 768 * 2 + 1152 * 2)
    assert ctypes.sizeof(L.TrunkLayer) == 56 and ctypes.sizeof(L.TrunkArgs) % 8 == 0
    assert lib.splade_trunk_encode(None, None) == 1               # SPLADE_ERR_INVALID_ARG, no crash
    assert b"null pointer" in lib.splade_last_error()
