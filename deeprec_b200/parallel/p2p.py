"""NVLink peer-memory communicator: CUDA-IPC symmetric buffers + the fused P2P kernels.

torch.distributed (NCCL) is used ONLY as plumbing: to exchange the 64-byte IPC handles at start-up.
On the step's hot path there is no NCCL call: csrc/cuda/comm_kernels.cu loads/stores peer memory
from inside the kernels and synchronises ranks with release/acquire flags over NVLink.

Per step (4 flag barriers, channel numbers in brackets):
  [0] ids landed on every rank      -> k_mp_lookup   (peer id loads, probe, peer row stores)
  [1] every owner finished writing  -> interaction / top MLP read the local receive buffer
  [2] gradient columns written      -> k_mp_sparse_grad (peer grad loads, dedup) + k_apply
  [3] dense gradients complete      -> k_allreduce_apply (peer grad loads, fixed order, + optimizer)
Buffer-reuse safety follows from barrier [0] of the next step.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch
import torch.distributed as dist

from .. import _native
from .._native import ptr

vp = C.c_void_p


class Peers(C.Structure):
    """Mirror of DrPeers (csrc/cuda/comm_kernels.cu)."""
    _fields_ = [("ptr", vp * 16)]


class _RawCuda:
    def __init__(self, addr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (addr, False), "version": 3}


def _bind(lib):
    if getattr(lib, "_comm_bound", False):
        return
    i64, INT, P = C.c_int64, C.c_int, C.c_void_p
    PP = C.POINTER(Peers)
    sigs = {
        "dr_cuda_set_device": [INT], "dr_cuda_get_device": [],
        "dr_comm_alloc": [i64, C.POINTER(vp)], "dr_comm_free": [P], "dr_comm_get_handle": [P, P],
        "dr_comm_open_handle": [P, C.POINTER(vp)], "dr_comm_close_handle": [P], "dr_comm_can_access_peer": [INT, INT],
        "dr_comm_barrier": [PP, P, INT, INT, INT, P],
        "dr_comm_mp_partition": [P, P, INT, INT, i64, P, P, P, P],
        "dr_comm_mp_lookup": [P, P, P, INT, INT, INT, INT, i64, INT, INT, PP, PP, PP, PP, PP, P, P, INT, P, P, P, P, i64, P],
        "dr_comm_mp_sparse_grad": [P, P, P, INT, INT, INT, i64, INT, PP, P, P, P, P, P],
        "dr_comm_allreduce_apply": [PP, INT, P, P, P, i64, P, P, P],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, INT
    lib._comm_bound = True


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"deeprec_cuda comm: {what} failed with code {rc}")


class SymmetricBuffer:
    """Same-size device allocation on every rank, mapped into every rank's address space."""

    def __init__(self, comm: "P2PComm", nbytes: int):
        lib = comm.lib
        self.comm, self.nbytes = comm, nbytes
        local = vp()
        _chk(lib.dr_comm_alloc(nbytes, C.byref(local)), "alloc")
        self.local = local.value
        handle = (C.c_char * 64)()
        _chk(lib.dr_comm_get_handle(vp(self.local), handle), "ipc_get_handle")
        handles: List[bytes] = [None] * comm.world
        dist.all_gather_object(handles, bytes(handle.raw), group=comm.group)
        self.peers = Peers()
        self._opened = []
        for r in range(comm.world):
            if r == comm.rank:
                self.peers.ptr[r] = self.local
            else:
                p = vp()
                buf = C.create_string_buffer(handles[r], 64)
                _chk(lib.dr_comm_open_handle(buf, C.byref(p)), f"ipc_open_handle(rank {r})")
                self.peers.ptr[r] = p.value
                self._opened.append(p.value)
        self._raw = torch.as_tensor(_RawCuda(self.local, nbytes), device=comm.dev)

    def tensor(self, dtype: torch.dtype, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nb = n * torch.empty((), dtype=dtype).element_size()
        assert nb <= self.nbytes
        return self._raw[:nb].view(dtype).view(*shape)

    def peers_ref(self):
        return C.byref(self.peers)


class P2PComm:
    supports_row_sharding = True

    def __init__(self, rank: int, world: int, dev: torch.device, group=None):
        self.rank, self.world, self.dev, self.group = rank, world, dev, group
        self.lib = _native.cuda()
        _bind(self.lib)
        _chk(self.lib.dr_cuda_set_device(dev.index), "set_device")
        for r in range(world):
            if r != rank and not self.lib.dr_comm_can_access_peer(dev.index, r):
                raise RuntimeError(f"GPU {dev.index} cannot access peer {r}: NVLink P2P is required")
        self.signals = SymmetricBuffer(self, 16 * 16 * 4)
        self.epochs = torch.zeros(16, dtype=torch.int32, device=dev)
        dist.barrier(group=group)
        self.launches = 0
        import os
        self._timing = os.environ.get("DEEPREC_P2P_TIMING") == "1"
        self._events = {}

    def _tick(self, name):
        if self._timing:
            e = torch.cuda.Event(enable_timing=True); e.record(torch.cuda.current_stream(self.dev))
            self._events.setdefault(name, []).append(e)

    def timing_report(self):
        torch.cuda.synchronize(self.dev)
        ev, out = self._events, {}
        names = [("barrier0", "l0", "l1"), ("mp_lookup", "l1", "l2"), ("barrier1", "l2", "l3"), ("barrier2", "s0", "s1"), ("mp_sparse_grad", "s1", "s2"),
                 ("sparse_apply", "s2", "s3"), ("barrier3", "d0", "d1"), ("allreduce_apply", "d1", "d2")]
        for nm, a, b in names:
            if a in ev and b in ev:
                ts = [x.elapsed_time(y) for x, y in zip(ev[a], ev[b])]
                out[nm] = sum(ts[2:]) / max(1, len(ts[2:]))
        return out

    def _s(self):
        return vp(torch.cuda.current_stream(self.dev).cuda_stream)

    # ---- buffers the engine asks for --------------------------------------------------------------------
    def alloc_grads(self, P: int) -> torch.Tensor:
        self.grads_buf = SymmetricBuffer(self, P * 4)
        return self.grads_buf.tensor(torch.float32, (P,))

    def alloc_exchange(self, T: int, B: int, D: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        self.ids_buf = SymmetricBuffer(self, T * B * 8)
        self.emb_buf = SymmetricBuffer(self, T * B * D * 2)
        self.demb_buf = SymmetricBuffer(self, T * B * D * 2)
        dist.barrier(group=self.group)
        return (self.ids_buf.tensor(torch.int64, (T, B)), self.emb_buf.tensor(torch.bfloat16, (T, B, D)),
                self.demb_buf.tensor(torch.bfloat16, (T, B, D)))

    def alloc_row_dispatch(self, nr: int, B: int) -> None:
        """Requester-side buckets for the row-sharded tables (ids grouped by owning rank, read by the owners over NVLink)."""
        self.nr = nr
        if nr == 0:
            return
        W = self.world
        self.bkt_key = SymmetricBuffer(self, nr * W * B * 8)
        self.bkt_b = SymmetricBuffer(self, nr * W * B * 4)
        self.bkt_cnt = SymmetricBuffer(self, max(256, nr * W) * 4)
        self.cnt_local = torch.zeros(nr * W, dtype=torch.int32, device=self.dev)
        self.brow = torch.zeros(nr * W * B, dtype=torch.int32, device=self.dev)
        dist.barrier(group=self.group)

    def barrier(self, channel: int) -> None:
        _chk(self.lib.dr_comm_barrier(self.signals.peers_ref(), ptr(self.epochs), channel, self.rank, self.world, self._s()), "barrier")

    # ---- the three fused paths ----------------------------------------------------------------------------
    def _tg(self, eng):
        if not hasattr(eng, "_table_global"):
            eng._table_global = torch.tensor(eng.local_tables, dtype=torch.int32, device=self.dev)
        return eng._table_global

    def lookup_forward(self, eng, train: bool) -> None:
        nl, nr, ctx = eng.n_tablewise, len(eng.row_tables), eng.ctx
        if not hasattr(self, "nr"):
            self.alloc_row_dispatch(nr, eng.B)
        if nr:
            # dispatch: bucket my ids of the row-sharded tables by owner before everybody meets at barrier 0
            _chk(self.lib.dr_comm_mp_partition(ptr(eng.ids), ptr(eng.row_tg), nr, self.world, eng.B, vp(self.bkt_key.local), vp(self.bkt_b.local),
                                               vp(self.bkt_cnt.local), self._s()), "mp_partition")
            eng.launches += 1
        self._tick("l0")
        self.barrier(0)
        self._tick("l1")
        _chk(self.lib.dr_comm_mp_lookup(ptr(ctx.structs()), ptr(eng.tmap_local), ptr(self._tg(eng)), self.rank, nl, nr, self.world, eng.B, eng.T, eng.D,
                                        self.ids_buf.peers_ref(), self.emb_buf.peers_ref(),
                                        self.bkt_key.peers_ref() if nr else None, self.bkt_b.peers_ref() if nr else None,
                                        self.bkt_cnt.peers_ref() if nr else None, ptr(self.cnt_local) if nr else None, ptr(self.brow) if nr else None,
                                        int(train), eng.step_ptr, ptr(eng.pos),
                                        ptr(ctx.ulist) if train else None, ptr(ctx.nuniq) if train else None,
                                        ctx.ulist.numel() if train else 0, self._s()), "mp_lookup")
        self._tick("l2")
        self.barrier(1)
        self._tick("l3")
        eng.launches += 3

    def sparse_backward(self, eng) -> None:
        nl, nr, ctx = eng.n_tablewise, len(eng.row_tables), eng.ctx
        self._tick("s0")
        self.barrier(2)
        self._tick("s1")
        _chk(self.lib.dr_comm_mp_sparse_grad(ptr(ctx.structs()), ptr(eng.tmap_local), ptr(self._tg(eng)), nl, nr, self.world, eng.B, eng.D,
                                             self.demb_buf.peers_ref(), ptr(eng.pos), ptr(self.cnt_local) if nr else None,
                                             ptr(self.brow) if nr else None, ptr(ctx.gsum), self._s()), "mp_sparse_grad")
        self._tick("s2")
        _chk(self.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), eng.D,
                                           ptr(eng.hp_dev), eng.max_unique, 1, self._s()), "sparse_apply")
        self._tick("s3")
        eng.launches += 4

    def dense_allreduce_update(self, eng) -> None:
        self._tick("d0")
        self.barrier(3)
        self._tick("d1")
        _chk(self.lib.dr_comm_allreduce_apply(self.grads_buf.peers_ref(), self.world, ptr(eng.params), ptr(eng.s0) if eng.s0 is not None else None,
                                              ptr(eng.s1) if eng.s1 is not None else None, eng.P, ptr(eng.hp_dev), None, self._s()), "allreduce_apply")
        self._tick("d2")
        eng.launches += 2

    def allreduce(self, out: torch.Tensor) -> None:
        """Plain one-shot all-reduce of the symmetric grads buffer into ``out`` (tests / metrics)."""
        self.barrier(3)
        _chk(self.lib.dr_comm_allreduce_apply(self.grads_buf.peers_ref(), self.world, None, None, None, out.numel(), None, ptr(out), self._s()), "allreduce")
