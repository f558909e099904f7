"""ctypes binding of ``libsplade_b200.so`` (C-ABI declared in ``include/splade_b200.h``).

The library is built in-tree by ``csrc/build.sh`` (see ``__graft_entry__.build``).  There is NO fallback:
if the shared object is missing or a call returns a non-zero code, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32, c_void_p)

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libsplade_b200.so")
ABI_VERSION = 5

# op ids for splade_workspace_bytes (include/splade_b200.h)
OP_HEAD_FWD = 0
OP_HEAD_FWD_TOKW = 1
OP_HEAD_BWD = 2
OP_FLOPS_NNZ = 3
OP_SCORES = 4
OP_V33_LOSS_FWD = 5
OP_TRUNK_ENCODE = 6
OP_TRUNK_GEMM = 7
OP_TRUNK_ATTN = 8
OP_PEER_LOSS_BWD = 9
OP_HEAD_BWD_DHIDDEN = 10
OP_HEAD_BWD_BITMAP = 11
OP_LOSS_BWD_TC = 12
OP_HEAD_PROLOGUE = 13
OP_TRUNK_TRAIN_FWD = 14
OP_TRUNK_LN_BWD = 15
HEAD_BWD_MAX_SEGMENTS = 8
MAX_PEERS = 16

# epilogues of splade_trunk_gemm
EPI_BF16, EPI_GELU, EPI_RESID, EPI_GEGLU = 0, 1, 2, 3


class V33ScalarArgs(Structure):
    """Mirror of ``splade_v33_scalar_args`` (include/splade_b200.h)."""

    _fields_ = [
        ("Bq", c_int32), ("N", c_int32), ("k", c_int32), ("label_offset", c_int32),
        ("round_bf16", c_int32), ("has_kd", c_int32), ("has_mmse", c_int32), ("reserved", c_int32),
        ("temperature", c_float), ("kd_temperature", c_float),
        ("lambda_q", c_float), ("lambda_d", c_float), ("lambda_neg", c_float),
        ("lambda_kd", c_float), ("lambda_mmse", c_float), ("reserved_f", c_float),
        ("raw_inbatch", c_void_p), ("raw_hard", c_void_p), ("raw_pos", c_void_p),
        ("teacher_kd", c_void_p), ("teacher_pos", c_void_p), ("teacher_neg", c_void_p),
        ("flops3", c_void_p),
        ("scalars", c_void_p), ("g_inbatch", c_void_p), ("g_hard", c_void_p), ("row_scratch", c_void_p),
    ]


class V33FwdArgs(Structure):
    """Mirror of ``splade_v33_fwd_args`` (include/splade_b200.h)."""

    _fields_ = [
        ("Bq", c_int32), ("N", c_int32), ("k", c_int32), ("V", c_int32),
        ("label_offset", c_int32), ("round_bf16", c_int32), ("has_kd", c_int32), ("has_mmse", c_int32),
        ("temperature", c_float), ("kd_temperature", c_float),
        ("lambda_q", c_float), ("lambda_d", c_float), ("lambda_neg", c_float),
        ("lambda_kd", c_float), ("lambda_mmse", c_float), ("reserved_f", c_float),
        ("q", c_void_p), ("docs", c_void_p), ("negs", c_void_p),
        ("teacher_kd", c_void_p), ("teacher_pos", c_void_p), ("teacher_neg", c_void_p),
        ("colmean", c_void_p), ("scalars", c_void_p), ("g_inbatch", c_void_p), ("g_hard", c_void_p),
        ("workspace", c_void_p), ("ws_bytes", c_size_t),
    ]


class HeadBwdSegment(Structure):
    """Mirror of ``splade_head_bwd_segment`` (include/splade_b200.h)."""

    _fields_ = [("g_rep", c_void_p), ("rep", c_void_p), ("argmax", c_void_p), ("hidden", c_void_p),
                ("d_hidden", c_void_p), ("B", c_int32), ("S", c_int32), ("active", c_void_p)]


class TrunkLayer(Structure):
    """Mirror of ``splade_trunk_layer`` (include/splade_b200.h)."""

    _fields_ = [
        ("attn_norm", c_void_p), ("wqkv", c_void_p), ("wo", c_void_p), ("mlp_norm", c_void_p),
        ("wi", c_void_p), ("wo2", c_void_p), ("is_global", c_int32), ("reserved", c_int32),
    ]


class TrunkArgs(Structure):
    """Mirror of ``splade_trunk_args`` (include/splade_b200.h)."""

    _fields_ = [
        ("B", c_int32), ("S", c_int32), ("H", c_int32), ("I", c_int32), ("V", c_int32),
        ("n_heads", c_int32), ("n_layers", c_int32), ("local_window", c_int32),
        ("norm_eps", c_float), ("global_theta", c_float), ("local_theta", c_float), ("reserved_f", c_float),
        ("input_ids", c_void_p), ("mask", c_void_p), ("tok_emb", c_void_p), ("emb_norm", c_void_p),
        ("layers", POINTER(TrunkLayer)),
        ("final_norm", c_void_p), ("head_dense", c_void_p), ("head_norm", c_void_p),
        ("dec_w", c_void_p), ("dec_bias", c_void_p), ("rep", c_void_p), ("argmax", c_void_p),
        ("hidden_out", c_void_p), ("meta_out", c_void_p), ("plan_out", c_void_p),
        ("workspace", c_void_p), ("ws_bytes", c_size_t),
    ]


class TrunkTrainSaves(Structure):
    """Mirror of ``splade_trunk_train_saves`` (include/splade_b200.h)."""

    _fields_ = [("x_attn_in", c_void_p), ("x_mlp_in", c_void_p), ("xb", c_void_p), ("qkv", c_void_p),
                ("ab", c_void_p), ("xb2", c_void_p), ("gg", c_void_p), ("x_final", c_void_p), ("out", c_void_p),
                ("meta", c_void_p), ("plan", c_void_p)]


class PeerGroup(Structure):
    """Mirror of ``splade_peer_group`` (include/splade_b200.h)."""

    _fields_ = [("rank", c_int32), ("world", c_int32), ("Bq", c_int32), ("V", c_int32),
                ("window", c_void_p * MAX_PEERS)]


# name -> (restype, argtypes); every symbol include/splade_b200.h declares
EXPORTS = {
    "splade_abi_version": (c_int, []),
    "splade_last_error": (c_char_p, []),
    "splade_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "splade_head_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_head_fwd_mode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p]),
    "splade_head_fwd_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                     c_void_p]),
    "splade_head_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_rep_bitmap": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "splade_head_bwd_multi_workspace_bytes": (c_size_t, [POINTER(HeadBwdSegment), c_int, c_int]),
    "splade_head_bwd_multi": (c_int, [POINTER(HeadBwdSegment), c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                      c_int, c_void_p, c_size_t, c_void_p]),
    "splade_head_bwd_dhidden": (c_int, [POINTER(HeadBwdSegment), c_void_p, c_int, c_int, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "splade_head_bwd_last": (c_int, [POINTER(HeadBwdSegment), POINTER(HeadBwdSegment), POINTER(c_void_p), c_int,
                                     c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                     c_void_p]),
    "splade_head_bwd_dw": (c_int, [POINTER(HeadBwdSegment), POINTER(c_void_p), c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_int, c_void_p]),
    "splade_flops_nnz": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_scores": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                              c_void_p]),
    "splade_rowdot": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "splade_v33_scalars": (c_int, [POINTER(V33ScalarArgs), c_void_p]),
    "splade_v33_loss_fwd": (c_int, [POINTER(V33FwdArgs), c_void_p]),
    "splade_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "splade_v33_loss_fwd_operands": (c_int, [POINTER(V33FwdArgs), POINTER(c_void_p), POINTER(c_void_p),
                                             POINTER(c_int)]),
    "splade_loss_bwd_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_emit_set_threads": (c_int, [c_int]),
    "splade_emit_json": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_size_t, c_void_p, POINTER(c_size_t)]),
    "splade_emit_bulk": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                 c_char_p, c_char_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p, c_size_t, c_void_p, POINTER(c_size_t)]),
    "splade_emit_search": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_char_p, c_char_p, c_void_p, c_size_t, c_void_p, POINTER(c_size_t)]),
    "splade_trunk_encode": (c_int, [POINTER(TrunkArgs), c_void_p]),
    "splade_packed_rows": (c_size_t, [c_int, c_int]),
    "splade_trunk_train_fwd": (c_int, [POINTER(TrunkArgs), POINTER(TrunkTrainSaves), c_void_p]),
    "splade_trunk_ln_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "splade_trunk_geglu_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "splade_trunk_gather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p]),
    "splade_trunk_dqkv_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "splade_trunk_gemm": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                                  c_size_t, c_void_p]),
    "splade_head_prologue_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_head_prologue_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_trunk_attn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "splade_topk": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "splade_peer_window_bytes": (c_size_t, [c_int, c_int, c_int]),
    "splade_peer_alloc": (c_int, [c_size_t, POINTER(c_void_p)]),
    "splade_peer_free": (c_int, [c_void_p]),
    "splade_peer_export": (c_int, [c_void_p, c_void_p]),
    "splade_peer_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "splade_peer_close": (c_int, [c_void_p]),
    "splade_peer_status": (c_int, [POINTER(PeerGroup), POINTER(c_int32)]),
    "splade_peer_v33_loss_fwd": (c_int, [POINTER(V33FwdArgs), POINTER(PeerGroup), c_uint32, c_void_p, c_void_p]),
    "splade_peer_loss_bwd": (c_int, [POINTER(PeerGroup), c_uint32, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load_library() -> ctypes.CDLL:
    """Load (once) and type the shared library.  Raises if it is missing or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"splade_b200: native library not found at {LIB_PATH}. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` (or csrc/build.sh). "
                f"There is no CPU / eager fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.splade_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"splade_b200: ABI version mismatch (library {got}, python {ABI_VERSION})")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().splade_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
