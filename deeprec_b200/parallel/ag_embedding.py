"""All-gather -> lookup -> reduce-scatter model-parallel embedding over peer memory (``csrc/cuda/ag_embedding.cu``) -- the fused-kernel form of SOK
v1's ``DistributedEmbedding`` dataflow (3 x ncclAllGather of the ids, local lookup + partial combine, ncclReduceScatter of the
partial sums, ncclAllGather of the top gradients in the backward).  No gathered copy of the ids exists: owners read every peer's list in place,
requesters pull and sum the owners' partial rows, owners pull the top gradients -- all synchronised by the in-kernel flags of ``sp_sync.cuh``.

One :class:`AllGatherEmbedding` per (table, rank).  Per step::

    ag.load_ids(values, row_ids)        # this rank's sparse batch: key + sample index per entry
    ag.lookup(ctx, train=True)          # owner side: every rank's entries that hash to me
    out = ag.reduce(combiner="mean")    # requester side: [B, D] combined embeddings of MY samples
    ...dense forward / backward...
    ag.stage_grad(dout)                 # top gradients of my samples
    ag.grad(ctx); ctx.apply(...)        # owner side: per-key gradient sums -> row optimizer
    ag.step_end()

``parallel/sok.py::DistributedEmbedding`` keeps the module-level NCCL form (any device / backend); the engines use the unique-first pipeline
(``parallel/sparse_pipeline.py``), which moves strictly less data for one-hot columns."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _native
from .._native import ptr
from .sparse_pipeline import Peers, SpSync, _LocalBuffer, bind as _bind_sp

vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
CH_IDS, CH_ROWS, CH_GRAD, CH_DONE = 0, 1, 2, 4


class AgGeom(C.Structure):
    """Mirror of DrAgGeom."""
    _fields_ = [("W", i32), ("rank", i32), ("dim", i32), ("table_index", i32), ("B", i64), ("nnz_cap", i64)]


class AgPeers(C.Structure):
    """Mirror of DrAgPeers."""
    _fields_ = [("keys", Peers), ("rows", Peers), ("meta", Peers), ("partial", Peers), ("grad", Peers)]


def _bind(lib):
    if getattr(lib, "_ag_bound", False):
        return lib
    INT, P = C.c_int, vp
    GP, PP, SP = C.POINTER(AgGeom), C.POINTER(AgPeers), C.POINTER(SpSync)
    lib.dr_ag_lookup.argtypes, lib.dr_ag_lookup.restype = [P, GP, PP, INT, P, P, P, P, P, i64, SP, P], INT
    lib.dr_ag_reduce.argtypes, lib.dr_ag_reduce.restype = [GP, PP, P, P, SP, P], INT
    lib.dr_ag_grad.argtypes, lib.dr_ag_grad.restype = [P, GP, PP, P, P, P, P, i64, SP, P], INT
    lib.dr_ag_sizeof_geom.argtypes, lib.dr_ag_sizeof_geom.restype = [], INT
    assert lib.dr_ag_sizeof_geom() == C.sizeof(AgGeom)
    lib._ag_bound = True
    return lib


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"deeprec_cuda all-gather embedding: {what} failed with code {rc}")


class AllGatherEmbedding:
    def __init__(self, dev: torch.device, rank: int, world: int, table, batch: int, nnz_cap: int, comm=None):
        """``table``: this rank's shard (a :class:`ops.device_table.DeviceTable` registered in a StepContext); ``nnz_cap``: most entries a rank's
        batch may have; ``comm``: :class:`parallel.p2p.P2PComm` (or the emulation's ``EmuComm``) when ``world > 1``."""
        if world > 1 and comm is None:
            raise ValueError("world > 1 needs a communicator for the symmetric buffers")
        if world > 16:
            raise ValueError("one peer-memory domain: at most 16 ranks")
        self.dev, self.rank, self.W, self.table, self.B, self.D, self.nnz_cap = dev, rank, world, table, int(batch), int(table.dim), int(nnz_cap)
        self.lib = _bind(_bind_sp(_native.cuda()))
        sym = (lambda nb: comm.symmetric(nb)) if world > 1 else (lambda nb: _LocalBuffer(dev, nb))
        self.keys_buf, self.rows_buf, self.meta_buf = sym(self.nnz_cap * 8), sym(self.nnz_cap * 4), sym(64)
        self.partial_buf, self.grad_buf, self.flags_buf = sym(world * self.B * self.D * 4), sym(self.B * self.D * 4), sym(8 * 16 * 4)
        self.keys = self.keys_buf.tensor(torch.int64, (self.nnz_cap,))
        self.rows = self.rows_buf.tensor(torch.int32, (self.nnz_cap,))
        self.meta = self.meta_buf.tensor(torch.int32, (4,))
        self.grad_stage = self.grad_buf.tensor(torch.float32, (self.B, self.D))
        p = AgPeers()
        p.keys, p.rows, p.meta, p.partial, p.grad = self.keys_buf.peers, self.rows_buf.peers, self.meta_buf.peers, self.partial_buf.peers, self.grad_buf.peers
        self.peers = p
        g = AgGeom()
        g.W, g.rank, g.dim, g.table_index, g.B, g.nnz_cap = world, rank, self.D, int(table.gid), self.B, self.nnz_cap
        self.geom = g
        self.state = torch.zeros(16, dtype=torch.int32, device=dev)
        s = SpSync()
        s.flags, s.state, s.rank, s.W = self.flags_buf.peers, self.state.data_ptr(), rank, world
        self.sync = s
        self.own_pos = torch.zeros(world * self.nnz_cap, dtype=torch.int32, device=dev)
        self.own_row = torch.zeros(world * self.nnz_cap, dtype=torch.int32, device=dev)
        self.own_cnt = torch.zeros(world, dtype=torch.int32, device=dev)
        self._scale: Optional[torch.Tensor] = None
        if world > 1:
            comm.host_barrier()

    def _s(self):
        return None if _native.emu_active() else vp(torch.cuda.current_stream(self.dev).cuda_stream)

    def _signal(self, ch: int) -> None:
        _chk(self.lib.dr_sp_signal(C.byref(self.sync), ch, self._s()), "signal")

    # ---- requester side -------------------------------------------------------------------------------------------------------------------
    def load_ids(self, values: torch.Tensor, row_ids: torch.Tensor) -> None:
        """Stage this rank's sparse batch (``values`` int64 keys, ``row_ids`` sample index of every entry) where the owners read it."""
        n = int(values.numel())
        if n > self.nnz_cap:
            raise ValueError(f"{n} entries > nnz_cap {self.nnz_cap}")
        self.keys[:n].copy_(values.reshape(-1)); self.rows[:n].copy_(row_ids.reshape(-1).to(torch.int32))
        self.meta[0] = n
        self._rows_now = row_ids.reshape(-1)
        self._signal(CH_IDS)

    def reduce(self, combiner: str = "sum") -> torch.Tensor:
        """[B, D] fp32: for every sample of this rank the combined rows of its entries (pulled from all owners' partial sums)."""
        self._scale = None
        if combiner != "sum":
            cnt = torch.bincount(self._rows_now.to(torch.int64), minlength=self.B).clamp(min=1).to(torch.float32)
            self._scale = (1.0 / cnt if combiner == "mean" else cnt.rsqrt()).to(self.dev).contiguous()
        out = torch.empty(self.B, self.D, dtype=torch.float32, device=self.dev)
        _chk(self.lib.dr_ag_reduce(C.byref(self.geom), C.byref(self.peers), ptr(self._scale), ptr(out), C.byref(self.sync), self._s()), "reduce")
        return out

    def stage_grad(self, dout: torch.Tensor) -> None:
        """Top gradients [B, D] of this rank's samples (scaled by the combiner here, so that owners add them as they are)."""
        g = dout.to(torch.float32)
        if self._scale is not None:
            g = g * self._scale.unsqueeze(1)
        self.grad_stage.copy_(g)
        self._signal(CH_GRAD)

    # ---- owner side -----------------------------------------------------------------------------------------------------------------------
    def lookup(self, ctx, train: bool) -> None:
        if train:
            ctx.ensure(self.W * self.nnz_cap)
        _chk(self.lib.dr_ag_lookup(ptr(ctx.structs()), C.byref(self.geom), C.byref(self.peers), int(train), ptr(self.own_pos), ptr(self.own_row), ptr(self.own_cnt),
                                   ptr(ctx.ulist) if train else None, ptr(ctx.nuniq) if train else None, ctx.ulist.numel() if train else 0, C.byref(self.sync), self._s()),
             "lookup")

    def grad(self, ctx) -> None:
        _chk(self.lib.dr_ag_grad(ptr(ctx.structs()), C.byref(self.geom), C.byref(self.peers), ptr(self.own_pos), ptr(self.own_row), ptr(self.own_cnt), ptr(ctx.gsum),
                                 ctx.ulist.numel(), C.byref(self.sync), self._s()), "grad")

    def step_end(self) -> None:
        _chk(self.lib.dr_sp_step_end(ptr(self.state), self._s()), "step_end")
