"""SparseOperationKit-style API (addons/sparse_operation_kit/legacy/sparse_operation_kit: ``sok.Init``, ``DistributedEmbedding``,
``All2AllDenseEmbedding``, ``sok.optimizers.utils.split_embedding_variable_from_others``, ``sok.Saver``) on this framework.

Row-wise model parallelism: key k lives on rank ``row_owner(k) = mix(k) % world`` (all2all_input_dispatcher.cu:36,78 uses
``key % gpu_count``; we hash first so strided id spaces stay balanced).  Each rank owns one EmbeddingVariable shard.

  * ``All2AllDenseEmbedding``  (dense ids [B, slots]):  all2all ids -> local gather -> all2all vectors (C2/C3), backward
    all2all of the output gradients (C4) into the owner's sparse gradient.
  * ``DistributedEmbedding``   (sparse ids, combiner):  all-gather ids -> local lookup + partial combine -> reduce-scatter (C5/C6).

These layers run over torch.distributed (gloo on CPU, nccl on GPU) so they work for any model; the DLRM engine's hot path uses
the fused NVLink kernels in csrc/cuda/comm_kernels.cu instead (same dataflow, no NCCL call).
"""
from __future__ import annotations

import zlib

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import nn

from ..config import EmbeddingVariableOption
from ..embedding_variable import EmbeddingVariable, get_embedding_variable
from ..ops.embedding_ops import SparseIds

_M1, _M2, _SALT = -4658895280553007687, -7723592293110705685, 6616326155283851669     # splitmix64 constants / salt as signed int64


def _lsr(x: torch.Tensor, n: int) -> torch.Tensor:
    """logical shift right of int64 bit patterns"""
    return (x >> n) & ((1 << (64 - n)) - 1)


def row_owner(keys: torch.Tensor, world: int) -> torch.Tensor:
    """Bit-exact twin of ``row_owner`` in csrc/cuda/comm_kernels.cu (dr_mix64(key ^ salt) >> 33) % W, so this layer, the
    fused kernels and checkpoints agree on which rank owns a key."""
    x = keys.to(torch.int64) ^ _SALT
    x = (x ^ _lsr(x, 30)) * _M1
    x = (x ^ _lsr(x, 27)) * _M2
    x = x ^ _lsr(x, 31)
    return torch.remainder(_lsr(x, 33), world)


def Init(**kw) -> None:
    """``sok.Init``: make sure the process group exists (one process per GPU)."""
    from .collective import CollectiveStrategy
    CollectiveStrategy()


def _world() -> Tuple[int, int]:
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def _all_to_all_v(send: List[torch.Tensor]) -> List[torch.Tensor]:
    """Variable-size all-to-all of 1-D/2-D tensors (count exchange + payload), gloo- and nccl-compatible."""
    rank, W = _world()
    if W == 1:
        return [send[0]]
    dev = send[0].device
    counts = torch.tensor([t.shape[0] for t in send], dtype=torch.int64, device=dev)
    rc = torch.empty_like(counts)
    dist.all_to_all_single(rc, counts)
    tail = send[0].shape[1:]
    recv = [torch.empty((int(n),) + tuple(tail), dtype=send[0].dtype, device=dev) for n in rc.tolist()]
    if dist.get_backend() == "gloo":          # gloo has no list all_to_all: pairwise isend/irecv
        reqs = []
        for p in range(W):
            if p == rank:
                recv[p].copy_(send[p])
                continue
            reqs.append(dist.isend(send[p].contiguous(), p))
            reqs.append(dist.irecv(recv[p], p))
        for r in reqs:
            r.wait()
    else:
        dist.all_to_all(recv, [t.contiguous() for t in send])
    return recv


class _A2ADense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ev: EmbeddingVariable, ids: torch.Tensor):
        rank, W = _world()
        flat = ids.reshape(-1)
        own = row_owner(flat, W)
        order = torch.argsort(own, stable=True)
        sorted_ids = flat[order]
        counts = torch.bincount(own, minlength=W).tolist()
        recv_ids = _all_to_all_v(list(torch.split(sorted_ids, counts)))                 # C2: id dispatch
        rcounts = [t.numel() for t in recv_ids]
        local = torch.cat(recv_ids) if recv_ids else flat.new_empty(0)
        with torch.enable_grad():
            rows = ev.lookup(local)                                                   # owner-side probe + gather
        back = _all_to_all_v(list(torch.split(rows.detach(), rcounts)))                 # C3: vectors back to requesters
        out_sorted = torch.cat(back)
        out = torch.empty_like(out_sorted)
        out[order] = out_sorted
        ctx.rows, ctx.order, ctx.counts, ctx.rcounts = rows, order, counts, rcounts
        return out.view(*ids.shape, ev.embedding_dim)

    @staticmethod
    def backward(ctx, g):
        g2 = g.reshape(-1, g.shape[-1])[ctx.order]
        recv = _all_to_all_v(list(torch.split(g2.contiguous(), ctx.counts)))            # C4: sparse-gradient return
        grad_rows = torch.cat(recv)
        torch.autograd.backward([ctx.rows], [grad_rows.to(ctx.rows.device)])          # records the owner's sparse gradient
        return None, None, None


class All2AllDenseEmbedding(nn.Module):
    """``sok.All2AllDenseEmbedding(max_vocabulary_size_per_gpu, embedding_vec_size, slot_num, nnz_per_slot)``:
    inputs [B, slot_num, nnz_per_slot] (or any int64 shape) -> [..., embedding_vec_size]."""

    def __init__(self, embedding_vec_size: int, slot_num: int = 1, nnz_per_slot: int = 1, max_vocabulary_size_per_gpu: int = 0,
                 name: str = "sok_dense", ev_option: Optional[EmbeddingVariableOption] = None, device=None):
        super().__init__()
        rank, _ = _world()
        self.slot_num, self.nnz_per_slot = slot_num, nnz_per_slot
        self.ev = get_embedding_variable(f"{name}/shard_{rank}", embedding_vec_size, ev_option=ev_option, device=device, seed=zlib.crc32(name.encode()) & 0xFFFF)
        self._anchor = nn.Parameter(torch.zeros(0), requires_grad=True)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if not torch.is_grad_enabled():
            with torch.enable_grad():
                return _A2ADense.apply(self._anchor, self.ev, inputs).detach()
        return _A2ADense.apply(self._anchor, self.ev, inputs)

    def embedding_variables(self) -> List[EmbeddingVariable]:
        return [self.ev]


class _DistSparse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ev: EmbeddingVariable, values, row_ids, batch_size: int, combiner: str):
        rank, W = _world()
        # C5: all-gather (values, row ids, batch sizes) -- every rank sees the global batch
        if W > 1:
            objs = [None] * W
            dist.all_gather_object(objs, (values.cpu(), row_ids.cpu(), int(batch_size)))
        else:
            objs = [(values.cpu(), row_ids.cpu(), int(batch_size))]
        offs, tot = [], 0
        for _, _, b in objs:
            offs.append(tot); tot += b
        gv = torch.cat([o[0] for o in objs]); gr = torch.cat([o[1] + off for o, off in zip(objs, offs)])
        mine = row_owner(gv, W) == rank
        lv, lr = gv[mine], gr[mine]
        with torch.enable_grad():
            rows = ev.lookup(lv.to(ev.device))
        partial = torch.zeros(tot, ev.embedding_dim, dtype=rows.dtype, device=rows.device)
        partial.index_add_(0, lr.to(rows.device), rows.detach())
        nnz = torch.bincount(gr, minlength=tot).clamp_(min=1).to(partial.dtype).to(partial.device)
        scale = torch.ones_like(nnz) if combiner == "sum" else (1.0 / nnz if combiner == "mean" else nnz.rsqrt())
        # C6: reduce-scatter of the partial combines (rank r keeps its own batch slice)
        if W > 1:
            dist.all_reduce(partial)
        lo = offs[rank]
        out = (partial * scale.unsqueeze(1))[lo: lo + batch_size]
        ctx.rows, ctx.lr, ctx.scale, ctx.tot, ctx.lo, ctx.B = rows, lr, scale, tot, lo, batch_size
        return out.to(values.device)

    @staticmethod
    def backward(ctx, g):
        rank, W = _world()
        full = torch.zeros(ctx.tot, g.shape[1], dtype=g.dtype, device=ctx.rows.device)
        full[ctx.lo: ctx.lo + ctx.B] = g.to(full.device)
        if W > 1:
            dist.all_reduce(full)                                                      # bwd: all-gather of the top gradients
        grad_rows = (full * ctx.scale.unsqueeze(1))[ctx.lr.to(full.device)]
        torch.autograd.backward([ctx.rows], [grad_rows])
        return (None,) * 6


class DistributedEmbedding(nn.Module):
    """``sok.DistributedEmbedding(combiner, max_vocabulary_size_per_gpu, embedding_vec_size, slot_num, max_nnz)``:
    sparse inputs (SparseIds) -> combined [B, embedding_vec_size]; rows are sharded row-wise over all ranks."""

    def __init__(self, combiner: str, embedding_vec_size: int, slot_num: int = 1, max_nnz: int = 1, max_vocabulary_size_per_gpu: int = 0,
                 name: str = "sok_sparse", ev_option: Optional[EmbeddingVariableOption] = None, device=None):
        super().__init__()
        assert combiner in ("sum", "mean", "sqrtn")
        rank, _ = _world()
        self.combiner = combiner
        self.ev = get_embedding_variable(f"{name}/shard_{rank}", embedding_vec_size, ev_option=ev_option, device=device, seed=zlib.crc32(name.encode()) & 0xFFFF)
        self._anchor = nn.Parameter(torch.zeros(0), requires_grad=True)

    def forward(self, sp: SparseIds) -> torch.Tensor:
        return _DistSparse.apply(self._anchor, self.ev, sp.values, sp.row_ids, sp.batch_size, self.combiner)

    def embedding_variables(self) -> List[EmbeddingVariable]:
        return [self.ev]


class _A2AStatic(torch.autograd.Function):
    """Dispatch / gather / return over a STATIC table shard (``weight [rows_per_rank, D]``; id k lives on rank k % W at local row k // W)."""

    @staticmethod
    def forward(ctx, weight: torch.Tensor, ids: torch.Tensor):
        rank, W = _world()
        flat = ids.reshape(-1)
        own = torch.remainder(flat, W)
        order = torch.argsort(own, stable=True)
        counts = torch.bincount(own, minlength=W).tolist()
        recv_ids = _all_to_all_v(list(torch.split(flat[order], counts)))
        rcounts = [t.numel() for t in recv_ids]
        local = torch.div(torch.cat(recv_ids) if recv_ids else flat.new_empty(0), W, rounding_mode="floor").clamp_(0, weight.shape[0] - 1)
        back = _all_to_all_v(list(torch.split(weight.detach()[local], rcounts)))
        out_sorted = torch.cat(back)
        out = torch.empty_like(out_sorted)
        out[order] = out_sorted
        ctx.save_for_backward(local, order)
        ctx.counts, ctx.shape = counts, weight.shape
        return out.view(*ids.shape, weight.shape[1])

    @staticmethod
    def backward(ctx, g):
        local, order = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])[order]
        recv = _all_to_all_v(list(torch.split(g2.contiguous(), ctx.counts)))
        gw = torch.zeros(ctx.shape, dtype=g.dtype, device=g.device)
        gw.index_add_(0, local, torch.cat(recv))
        return gw, None


class TFDistributedEmbedding(nn.Module):
    """``sok.TFDistributedEmbedding(vocabulary_size, embedding_vec_size, initializer)``: the framework-native twin of the SOK layers -- a
    plain dense parameter of ``vocabulary_size`` rows, sharded row-wise over the ranks (id ``k`` on rank ``k % world`` at row ``k // world``),
    trained by the ordinary dense optimizer; no hash table, no admission.  The reference ships it as the correctness oracle of the custom
    layers (``sparse_operation_kit/embeddings/tf_distributed_embedding.py``); it plays the same role in ``tests/test_sok_elastic_cpu.py``."""

    def __init__(self, vocabulary_size: int, embedding_vec_size: int, initializer="uniform", device=None, dtype=torch.float32):
        super().__init__()
        _, W = _world()
        self.vocabulary_size, self.rows = int(vocabulary_size), (int(vocabulary_size) + W - 1) // W
        w = torch.empty(self.rows, embedding_vec_size, device=device, dtype=dtype)
        if callable(initializer):
            initializer(w)
        elif initializer in ("uniform", "random_uniform"):
            nn.init.uniform_(w, -0.05, 0.05)
        elif initializer == "ones":
            nn.init.ones_(w)
        elif initializer == "zeros":
            nn.init.zeros_(w)
        else:
            nn.init.normal_(w, std=0.01)
        self.weight = nn.Parameter(w)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        if (inputs < 0).any() or (inputs >= self.vocabulary_size).any():
            raise IndexError("TFDistributedEmbedding: id outside [0, vocabulary_size)")
        return _A2AStatic.apply(self.weight, inputs)


def split_embedding_variable_from_others(module: nn.Module):
    """``sok.optimizers.utils.split_embedding_variable_from_others``: (embedding variables, other parameters)."""
    from ..optim.optimizers import collect_embedding_variables
    evs = collect_embedding_variables(module)
    anchors = {id(e._anchor) for e in evs}
    return evs, [p for p in module.parameters() if id(p) not in anchors and p.numel() > 0]


class Saver:
    """``sok.Saver``: dump / restore the row-sharded tables (each rank writes its shard; restore re-shards by ownership)."""

    def dump_to_file(self, layer, path: str) -> None:
        rank, W = _world()
        snap = layer.ev.table.snapshot()
        torch.save({k: v.cpu() for k, v in snap.items()}, f"{path}.shard{rank}-of-{W}")
        if W > 1:
            dist.barrier()

    def restore_from_file(self, layer, path: str) -> int:
        import glob
        rank, W = _world()
        n = 0
        for f in sorted(glob.glob(f"{path}.shard*-of-*")):
            snap = torch.load(f)
            mine = row_owner(snap["keys"], W) == rank
            if mine.any():
                n += layer.ev.table.import_(snap["keys"][mine], snap["rows"][mine], snap["freqs"][mine], snap["versions"][mine])
        if W > 1:
            dist.barrier()
        return n


class _Utils:
    split_embedding_variable_from_others = staticmethod(split_embedding_variable_from_others)


class _Optimizers:
    """``sok.optimizers``: ``Adam`` / ``LazyAdam`` update the sharded tables row-wise (only rows that received a gradient move, their moments
    included -- SOK's ``update_functions.cu`` semantics, which is what this framework's sparse Adam does for every EmbeddingVariable);
    ``utils.split_embedding_variable_from_others`` separates the tables from the dense parameters of a model."""
    utils = _Utils

    @staticmethod
    def Adam(layers_or_evs, lr: float = 0.001, beta1: float = 0.9, beta2: float = 0.999, epsilon: float = 1e-8):
        from ..optim.optimizers import AdamOptimizer
        return AdamOptimizer([], _as_evs(layers_or_evs), lr=lr, beta1=beta1, beta2=beta2, eps=epsilon)

    @staticmethod
    def LazyAdam(layers_or_evs, lr: float = 0.001, beta1: float = 0.9, beta2: float = 0.999, epsilon: float = 1e-8):
        from ..optim.optimizers import AdamAsyncOptimizer
        return AdamAsyncOptimizer([], _as_evs(layers_or_evs), lr=lr, beta1=beta1, beta2=beta2, eps=epsilon)


def _as_evs(x) -> List[EmbeddingVariable]:
    items = x if isinstance(x, (list, tuple)) else [x]
    out: List[EmbeddingVariable] = []
    for it in items:
        out += it.embedding_variables() if hasattr(it, "embedding_variables") else [it]
    return out


optimizers = _Optimizers
