"""Communicator for the CUDA-on-CPU emulation (``_native.cuda_emulation()``): the ranks of a job are THREADS of one process, a symmetric buffer
is one host allocation per rank whose address every other rank simply knows -- the same ``symmetric()`` / ``host_barrier()`` surface as
:class:`parallel.p2p.P2PComm`, so :class:`parallel.sparse_pipeline.SparsePipeline` and the flag protocol of ``csrc/cuda/sp_sync.cuh`` run unchanged,
rank against rank, under ThreadSanitizer (tests/test_cuda_emu_sparse_pipeline.py).  CPU CI only."""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict, List

import torch

from .sparse_pipeline import Peers


class EmuWorld:
    """Shared by the rank threads: rendezvous barrier + the table of published buffer addresses."""

    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.lock = threading.Lock()
        self.published: Dict[int, List[int]] = {}


class _EmuSymmetricBuffer:
    def __init__(self, comm: "EmuComm", nbytes: int, index: int):
        nbytes = max(256, int(nbytes))
        self.nbytes = nbytes
        self._raw = torch.zeros(nbytes, dtype=torch.uint8)
        self.local = self._raw.data_ptr()
        w = comm.shared
        with w.lock:
            w.published.setdefault(index, [0] * w.world)[comm.rank] = self.local
        w.barrier.wait()
        self.peers = Peers()
        for r in range(w.world):
            self.peers.ptr[r] = w.published[index][r]
        w.barrier.wait()

    def tensor(self, dtype: torch.dtype, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert nb <= self.nbytes
        return self._raw[:nb].view(dtype).view(*shape)

    def peers_ref(self):
        return C.byref(self.peers)


class EmuComm:
    unique_first = True

    def __init__(self, shared: EmuWorld, rank: int):
        self.shared, self.rank, self.world = shared, rank, shared.world
        self.dev = torch.device("cpu")
        self._n = 0
        self._buffers = []

    def symmetric(self, nbytes: int) -> _EmuSymmetricBuffer:
        """Collective: every rank thread calls it in the same order."""
        b = _EmuSymmetricBuffer(self, nbytes, self._n)
        self._n += 1
        self._buffers.append(b)
        return b

    def host_barrier(self) -> None:
        self.shared.barrier.wait()
