"""Exchange windows for cross-GPU negatives over NVLink peer memory (``splade_peer_*`` in include/splade_b200.h).

``all_gather_documents`` (distributed.py) is the NCCL statement of the north-star exchange step: all-gather the
positives, score, reduce-scatter their gradients.  ``PeerExchange`` is the fused one: every rank owns a window
of device memory exported with CUDA IPC; the loss kernels pull the other ranks' positives (forward) and
queries (backward) straight out of the peers' windows, synchronised by flags written into those windows, so
neither the gathered gradient ``[W*B, V]`` nor an NCCL call is on the data path.  ``torch.distributed`` is used
once, at construction, to swap the 64-byte IPC handles.

Collective contract: every rank of the group constructs the exchange with the same (B, V), calls the loss the
same number of times and (when training) runs backward for each of them - the same rules as for an NCCL
collective.  A peer that never shows up trips the barrier timeout (``SPLADE_B200_PEER_TIMEOUT_MS``) and
``check()`` raises instead of the GPU hanging.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, load_library


class PeerExchange:
    """One rank's view of the W exchange windows."""

    def __init__(self, batch: int, vocab: int, group=None, device: Optional[torch.device] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group "
                               "(use PeerExchange.virtual_world for single-process tests)")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        lib = load_library()
        # Construction is collective and must FAIL collectively: a rank that cannot allocate, export or map a
        # window still takes part in both handle/status exchanges, then every rank raises the same error.
        err, handle = None, None
        try:
            self._setup(batch, vocab, rank, world, device)
            buf = ctypes.create_string_buffer(64)
            check(lib.splade_peer_export(self._own, buf), "splade_peer_export")
            handle = bytes(buf.raw)
        except Exception as e:  # noqa: BLE001
            err = f"rank {rank}: {e}"
        self._group = group
        published: List[Optional[tuple]] = [None] * world
        dist.all_gather_object(published, (handle, err), group=group)
        errors = [e for _, e in published if e]
        windows: List[int] = []
        if not errors:
            try:
                with torch.cuda.device(self.device):
                    for s in range(world):
                        if s == rank:
                            windows.append(self._own.value)
                            continue
                        p = ctypes.c_void_p()
                        check(lib.splade_peer_open(ctypes.create_string_buffer(published[s][0], 64),
                                                   ctypes.byref(p)), f"splade_peer_open(window of rank {s})")
                        self._opened.append(p)
                        windows.append(p.value)
            except Exception as e:  # noqa: BLE001
                err = f"rank {rank}: {e}"
            mapped: List[Optional[str]] = [None] * world
            dist.all_gather_object(mapped, err, group=group)  # also: nobody signals before every window is mapped
            errors = [e for e in mapped if e]
        if errors:
            self._release_local()
            raise RuntimeError("PeerExchange could not be set up: " + "; ".join(errors))
        self._bind(windows)
        self._distributed = True

    # -- construction helpers ----------------------------------------------------------------------------
    def _setup(self, batch: int, vocab: int, rank: int, world: int, device) -> None:
        lib = load_library()
        if world > _lib.MAX_PEERS:
            raise RuntimeError(f"PeerExchange supports at most {_lib.MAX_PEERS} ranks (got {world})")
        if vocab % 4 != 0:
            raise RuntimeError(f"PeerExchange needs a vocabulary size that is a multiple of 4 (got {vocab})")
        self.batch, self.vocab, self.rank, self.world = int(batch), int(vocab), int(rank), int(world)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.epoch = 0
        self._opened: List[ctypes.c_void_p] = []
        self._group = None
        self._distributed = False
        nbytes = lib.splade_peer_window_bytes(self.batch, self.vocab, self.world)
        if nbytes == 0:
            raise RuntimeError("splade_peer_window_bytes: bad shapes")
        self.window_bytes = nbytes
        self._own = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib.splade_peer_alloc(nbytes, ctypes.byref(self._own)), "splade_peer_alloc")

    def _bind(self, windows: List[int]) -> None:
        g = _lib.PeerGroup()
        g.rank, g.world, g.Bq, g.V = self.rank, self.world, self.batch, self.vocab
        for s, w in enumerate(windows):
            g.window[s] = w
        self.group_struct = g

    @classmethod
    def virtual_world(cls, batch: int, vocab: int, world: int, device=None) -> List["PeerExchange"]:
        """W exchanges in ONE process whose windows all live on `device` (tests: the same kernels and flags,
        same-device pointers instead of IPC mappings).  Drive every virtual rank on its own CUDA stream."""
        xs = []
        for r in range(world):
            x = cls.__new__(cls)
            x._setup(batch, vocab, r, world, device)
            xs.append(x)
        windows = [x._own.value for x in xs]
        for x in xs:
            x._bind(windows)
        return xs

    # -- per-step -----------------------------------------------------------------------------------------
    def next_epoch(self) -> int:
        if not (self._own and self._own.value):
            raise RuntimeError("PeerExchange used after close()")
        self.epoch += 1
        return self.epoch

    def check(self) -> None:
        """Raise if any barrier of this rank timed out (synchronises the device)."""
        if not (self._own and self._own.value):
            raise RuntimeError("PeerExchange.check() after close()")
        st = ctypes.c_int32(0)
        with torch.cuda.device(self.device):
            check(load_library().splade_peer_status(ctypes.byref(self.group_struct), ctypes.byref(st)),
                  "splade_peer_status")
        if st.value != 0:
            self.raise_for_status(st.value)

    def raise_for_status(self, status: int) -> None:
        """Turn a non-zero window status word (``splade_peer_status`` / scalars[15] of the peer loss) into the error."""
        phase, peer = status >> 8, (status & 0xFF) - 1
        raise RuntimeError(f"PeerExchange: rank {self.rank} waited in vain for rank {peer} "
                           f"({'forward' if phase == 1 else 'backward'} barrier timed out); "
                           f"the results of that step are invalid")

    def _release_local(self) -> None:
        lib = load_library()
        dev = getattr(self, "device", None)
        if dev is None:
            return
        with torch.cuda.device(dev):
            for p in getattr(self, "_opened", []):
                lib.splade_peer_close(p)
            self._opened = []
            if getattr(self, "_own", None) and self._own.value:
                lib.splade_peer_free(self._own)
                self._own = ctypes.c_void_p()

    def close(self) -> None:
        """Collective for a distributed exchange (every rank calls it): mappings are closed, then a barrier, then
        the owners free their windows."""
        lib = load_library()
        if not (self._own and self._own.value):
            return  # already closed
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.splade_peer_close(p)
            self._opened = []
            if self._distributed and dist.is_initialized():
                dist.barrier(group=self._group)
            self._release_local()

    def __del__(self):
        # close() is the documented way.  A distributed exchange is NOT torn down from a destructor: closing is
        # collective (barrier) and garbage collection is not; process exit releases the mappings.
        try:
            if not getattr(self, "_distributed", True) and getattr(self, "_own", None) and self._own.value:
                self.close()
        except Exception:
            pass
