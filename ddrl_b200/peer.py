"""Peer-mapped exchange buffers for the in-kernel gradient all-reduce (one process per GPU, one node).

Every rank allocates one exchange buffer through the C ABI (``ddrl_peer_alloc``: cudaMalloc + CUDA
IPC handle), the 64-byte handles travel through ``torch.distributed.all_gather_object`` and every rank maps the other
ranks' allocations (``ddrl_peer_open``).  The fused SGD tail (csrc/sgd_tail.cuh) then pushes its gradient slices straight
into the peers' buffers over NVLink as 64-bit {value, sequence} words; no NCCL call is made on the per-step path."""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import DDRLError, MAX_RANKS


class PeerExchange:
    """Exchange buffers sized for (P policies, NP parameters, G CTAs per policy) on ``world`` ranks."""

    def __init__(self, dist, world: int, rank: int, P: int, NP: int, G: int, device):
        if world > MAX_RANKS:
            raise DDRLError(f"in-kernel all-reduce supports up to {MAX_RANKS} ranks, got {world}")
        lib = _lib.load()
        self.world, self.rank, self.key = world, rank, (P, NP, G)
        words = int(lib.ddrl_sgd_exchange_words(NP, G))
        if words <= 0:
            raise DDRLError("ddrl_sgd_exchange_words failed")
        self.x_bytes = 2 * world * P * words * 8          # two step parities x {value, sequence} words
        self._own: List[int] = []
        self._opened: List[int] = []
        with torch.cuda.device(device):
            hx, px = self._alloc(lib, self.x_bytes)
            torch.cuda.synchronize()
            handles = [None] * world
            dist.all_gather_object(handles, hx)
            self.x_ptrs = [px if w == rank else self._open(lib, handles[w]) for w in range(world)]
            dist.barrier()     # everybody has mapped everybody before the first step may push
        self.seq = torch.zeros(1, dtype=torch.int32, device=device)

    def _alloc(self, lib, nbytes: int):
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(lib.ddrl_peer_alloc(nbytes, C.byref(ptr), handle), "peer_alloc")
        self._own.append(ptr.value)
        return handle.raw, ptr.value

    def _open(self, lib, handle: bytes) -> int:
        ptr = C.c_void_p()
        buf = C.create_string_buffer(handle, 64)
        _lib.check(lib.ddrl_peer_open(buf, C.byref(ptr)), "peer_open")
        self._opened.append(ptr.value)
        return ptr.value

    def fill(self, tail) -> None:
        """Write world / rank / seq / peer pointers into a ``SgdTail`` struct."""
        tail.world, tail.rank = self.world, self.rank
        tail.seq = self.seq.data_ptr()
        for w in range(self.world):
            tail.peer_x[w] = self.x_ptrs[w]

    def close(self) -> None:
        lib = _lib.load()
        for p in self._opened:
            lib.ddrl_peer_close(p)
        for p in self._own:
            lib.ddrl_peer_free(p)
        self._opened, self._own = [], []
