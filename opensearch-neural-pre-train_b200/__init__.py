"""splade_b200 — B200-native (sm_100a) SPLADE-max head + SPLADELossV33 hot path.

Drop-in for ``SPLADEModernBERT`` / ``SPLADELossV33`` of opensearch-neural-pre-train
(src/model/splade_modern.py, src/model/losses.py).  The arithmetic lives in ``libsplade_b200.so``
(hand-written tcgen05/TMEM/TMA CUDA, C-ABI in ``include/splade_b200.h``); this package is the host-side
mirror of the reference's Python interface.  There is no CPU fallback.
"""
from ._lib import ABI_VERSION, LIB_PATH, load_library  # noqa: F401
from .distributed import all_gather_documents, shard_range  # noqa: F401
from .encoder import SPLADEModernBERT  # noqa: F401
from .ops import (HostEncodeSlot, clear_weight_cache, flops_nnz, head_backward, head_backward_dhidden,  # noqa: F401
                  head_backward_dw, head_backward_multi, head_forward, head_forward_host, rep_bitmap, rowdot, scores,
                  splade_head, topk)
from .peer_exchange import PeerExchange  # noqa: F401
from .sparse_emit import SparseEmitter, build_blocklist, bulk_encode  # noqa: F401
from .trunk import NativeTrunk, trunk_attention, trunk_gemm  # noqa: F401
from .trunk_train import NativeTrunkTrain  # noqa: F401
from .v33_loss import SPLADELossV33  # noqa: F401

__all__ = [
    "SPLADEModernBERT", "SPLADELossV33", "PeerExchange", "SparseEmitter", "build_blocklist", "bulk_encode",
    "splade_head", "head_forward", "head_forward_host", "HostEncodeSlot", "head_backward", "head_backward_multi",
    "head_backward_dhidden", "head_backward_dw", "rep_bitmap", "clear_weight_cache", "flops_nnz", "scores", "rowdot", "topk",
    "NativeTrunk", "NativeTrunkTrain", "trunk_gemm", "trunk_attention", "all_gather_documents", "shard_range", "load_library", "LIB_PATH", "ABI_VERSION",
]
