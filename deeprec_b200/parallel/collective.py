"""CollectiveStrategy facade + GroupEmbedding strategy switch.

Parity: python/distribute/group_embedding_collective_strategy.py:28-138 (``CollectiveStrategy.{scope, embedding_scope,
world_size, rank, estimator, export_saved_model}``), python/framework/group_embedding_types.py:24-49 (strategy enum),
``tf.config.experimental.enable_distributed_strategy`` (framework/config.py:626), env ``COLLECTIVE_STRATEGY`` (sok | hb).
Here there is one built-in backend: model-parallel embedding shards over NVLink P2P (parallel/p2p.py) with the dense net
data-parallel; ``localized`` = single-process fused GroupEmbedding.
"""
from __future__ import annotations

import contextlib
import enum
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import strategy as _strategy


def _a2a(send: List[torch.Tensor]) -> List[torch.Tensor]:
    from .sok import _all_to_all_v
    return _all_to_all_v(send)


class _GroupDistLookup(torch.autograd.Function):
    """Model-parallel GroupEmbedding for generic modules, ALL tables of the call in one exchange (table t lives on rank t % world):

      forward   all-to-all #1  per-table (nnz, batch) of every requester          -> owners         (tiny)
                all-to-all #2  ids + row ids of every table, one message per owner                    (C2: id dispatch)
                owner: ONE lookup + combine per owned table over the concatenated requests of all ranks
                all-to-all #3  combined rows back, one message per requester                           (C3: embedding combine)
      backward  all-to-all #4  output gradients to the owners, whose EmbeddingVariables record the sparse gradients  (C4)

    Tensors only (no pickling, no host round trip): nccl on GPUs, gloo on CPU.  The flagship engine replaces these collectives by
    the fused NVLink kernels (parallel/p2p.py); this is the path every other model uses."""

    @staticmethod
    def forward(ctx, anchor, strategy, tables, sps, combiners, weights):
        from ..ops.embedding_ops import SparseIds, embedding_lookup_sparse
        W, r, T = strategy.world_size, strategy.rank, len(tables)
        dev = sps[0].values.device
        owned_by = [[t for t in range(T) if strategy.owner_of(t) == p] for p in range(W)]
        mine = owned_by[r]
        ws = [w if w is not None else sp.weights for w, sp in zip(weights, sps)]
        has_w = [w is not None for w in ws]
        # ---- #1 meta + #2 ids
        meta = [torch.tensor([[sps[t].values.numel(), sps[t].batch_size] for t in owned_by[p]], dtype=torch.int64, device=dev).reshape(-1, 2) for p in range(W)]
        ids = [torch.cat([torch.stack([sps[t].values.to(torch.int64), sps[t].row_ids.to(torch.int64)], 1) for t in owned_by[p]])
               if owned_by[p] else torch.zeros(0, 2, dtype=torch.int64, device=dev) for p in range(W)]
        rmeta, rids = _a2a(meta), _a2a(ids)
        rw = None
        if any(has_w):
            wsend = [torch.cat([(ws[t] if has_w[t] else torch.ones(sps[t].values.numel(), device=dev)).to(torch.float32) for t in owned_by[p]])
                     if owned_by[p] else torch.zeros(0, device=dev) for p in range(W)]
            rw = _a2a(wsend)
        # ---- owner: one lookup + combine per owned table over all requesters
        outs, out_batches, owner_sps = [], [], []
        cursor = [0] * W
        for k, t in enumerate(mine):
            vals, rows, wts, batches, base = [], [], [], [], 0
            for s_ in range(W):
                nnz, B = int(rmeta[s_][k, 0]), int(rmeta[s_][k, 1])
                seg = rids[s_][cursor[s_]: cursor[s_] + nnz]
                vals.append(seg[:, 0]); rows.append(seg[:, 1] + base)
                if rw is not None:
                    wts.append(rw[s_][cursor[s_]: cursor[s_] + nnz])
                cursor[s_] += nnz
                batches.append(B); base += B
            sp = SparseIds(torch.cat(vals), torch.cat(rows), base, torch.cat(wts) if (rw is not None and has_w[t]) else None)
            owner_sps.append(sp)
            out_batches.append(batches)
        from ..embedding_variable import EmbeddingVariable
        with torch.enable_grad():                                                     # [sum B, D_t] per owned table, autograd graph kept for backward
            if mine and all(isinstance(tables[t], EmbeddingVariable) and tables[t].device.type == "cuda" for t in mine):
                from ..ops.device_table import group_lookup_sparse_device             # ONE fused probe + gather + combine launch for all owned tables
                outs = list(group_lookup_sparse_device([tables[t] for t in mine], owner_sps, [combiners[t] for t in mine], [None] * len(mine)))
            else:
                outs = [embedding_lookup_sparse(tables[t], sp, None, combiners[t]) for t, sp in zip(mine, owner_sps)]
        # ---- #3 rows back to the requesters (flattened: tables may have different dims)
        back = []
        for s_ in range(W):
            parts, off = [], None
            for k in range(len(mine)):
                lo = sum(out_batches[k][:s_])
                parts.append(outs[k].detach()[lo: lo + out_batches[k][s_]].reshape(-1))
            back.append(torch.cat(parts) if parts else torch.zeros(0, device=dev))
        recv = _a2a([b.to(torch.float32) for b in back])
        results: List[Optional[torch.Tensor]] = [None] * T
        for p in range(W):
            off = 0
            for t in owned_by[p]:
                B, D = sps[t].batch_size, tables[t].embedding_dim
                results[t] = recv[p][off: off + B * D].view(B, D)
                off += B * D
        ctx.strategy, ctx.owned_by, ctx.outs, ctx.out_batches = strategy, owned_by, outs, out_batches
        ctx.dims = [tb.embedding_dim for tb in tables]
        ctx.batches = [sp.batch_size for sp in sps]
        return tuple(results)

    @staticmethod
    def backward(ctx, *grads):
        st, W, r = ctx.strategy, ctx.strategy.world_size, ctx.strategy.rank
        dev = grads[0].device
        send = [torch.cat([grads[t].reshape(-1).to(torch.float32) for t in ctx.owned_by[p]]) if ctx.owned_by[p] else torch.zeros(0, device=dev) for p in range(W)]
        recv = _a2a(send)                                                             # C4: gradients travel to the owners
        mine = ctx.owned_by[r]
        if mine:
            gouts, cursor = [], [0] * W
            for k, t in enumerate(mine):
                D, parts = ctx.dims[t], []
                for s_ in range(W):
                    n = ctx.out_batches[k][s_] * D
                    parts.append(recv[s_][cursor[s_]: cursor[s_] + n].view(-1, D)); cursor[s_] += n
                gouts.append(torch.cat(parts).to(ctx.outs[k].device))
            torch.autograd.backward(ctx.outs, gouts)                                  # the owner's EmbeddingVariables record the sparse gradients
        return (None,) * 6


class DistStrategy(enum.Enum):
    SOK = "sok"                 # accepted for API parity: maps onto the built-in P2P model-parallel path
    HB = "hb"
    COLLECTIVE = "collective"
    LOCALIZED = "localized"
    UNKNOWN = "unknown"


_ENABLED = DistStrategy.LOCALIZED


def enable_distributed_strategy(strategy: str = "collective") -> None:
    global _ENABLED
    _ENABLED = DistStrategy(strategy if strategy in ("sok", "hb", "collective", "localized") else "unknown")


class CollectiveStrategy:
    def __init__(self, backend: Optional[str] = None):
        name = os.environ.get("COLLECTIVE_STRATEGY", "collective")
        self.kind = DistStrategy(name) if name in ("sok", "hb", "collective", "localized") else DistStrategy.COLLECTIVE
        self.in_embedding_scope = False
        self._in_scope = False
        if dist.is_available() and not dist.is_initialized() and "RANK" in os.environ:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            be = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if be == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            dist.init_process_group(be)
        self._mp_tables = {}
        _strategy.set_current(self)

    @property
    def world_size(self) -> int:
        return dist.get_world_size() if dist.is_initialized() else 1

    @property
    def rank(self) -> int:
        return dist.get_rank() if dist.is_initialized() else 0

    @contextlib.contextmanager
    def scope(self):
        """Everything under the scope is data-parallel (dense parameters are kept bitwise in sync by the fused
        all-reduce + optimizer kernel; the hvd.BroadcastGlobalVariablesHook analogue is ``broadcast_parameters``)."""
        self._in_scope = True
        try:
            yield self
        finally:
            self._in_scope = False

    @contextlib.contextmanager
    def embedding_scope(self):
        """Lookups issued here are model-parallel: table t lives on rank t % world (table-wise, embedding_ops.py:1671-1676)."""
        self.in_embedding_scope = True
        try:
            yield self
        finally:
            self.in_embedding_scope = False

    def broadcast_parameters(self, module: torch.nn.Module, src: int = 0) -> None:
        if self.world_size > 1:
            for p in module.parameters():
                if p.numel():
                    dist.broadcast(p.data, src)

    def allreduce_gradients(self, params, bucket_bytes: int = 64 << 20, average: bool = False) -> None:
        """Dense gradient all-reduce for generic modules (the flagship engine fuses this with the optimizer): gradients are packed
        into flat buckets of <= ``bucket_bytes`` (Horovod's tensor fusion, 64 MB default) so a model with hundreds of small
        parameters costs a handful of collectives; ``average`` divides by the world size (Horovod's default op is Sum with the
        learning rate scaled instead -- see ``scale_learning_rate``)."""
        if self.world_size <= 1:
            return
        grads = [p.grad for p in params if p.grad is not None]
        by_kind = {}
        for g in grads:
            by_kind.setdefault((g.dtype, g.device), []).append(g)
        for (dtype, device), gs in by_kind.items():
            cap = max(1, bucket_bytes // max(1, gs[0].element_size()))
            i = 0
            while i < len(gs):
                j, n = i, 0
                while j < len(gs) and (n == 0 or n + gs[j].numel() <= cap):
                    n += gs[j].numel(); j += 1
                flat = torch.cat([g.reshape(-1) for g in gs[i:j]]) if j - i > 1 else gs[i].reshape(-1)
                dist.all_reduce(flat)
                if average:
                    flat /= self.world_size
                if j - i > 1:
                    o = 0
                    for g in gs[i:j]:
                        g.copy_(flat[o:o + g.numel()].view_as(g)); o += g.numel()
                elif not gs[i].is_contiguous():
                    gs[i].copy_(flat.view_as(gs[i]))
                i = j

    def scale_learning_rate(self, optimizer) -> None:
        """hvd_strategy.py:378-405: with Sum all-reduce the effective batch grows with the world size; the reference multiplies the
        learning rate of every wrapped optimizer by ``world_size``."""
        w = self.world_size
        if w > 1:
            for g in optimizer.param_groups:
                g["lr"] = g["lr"] * w
            if hasattr(optimizer, "lr"):
                optimizer.lr = optimizer.lr * w

    def estimator(self, model: torch.nn.Module, optimizer, loss_fn, checkpoint_dir: Optional[str] = None, **trainer_kw):
        """``CollectiveStrategy.estimator`` (group_embedding_collective_strategy.py:104-121): a training driver wired to this strategy --
        rank-0 parameter broadcast at start, bucketed gradient all-reduce every step, rank-suffixed checkpoints."""
        from ..utils.trainer import Trainer
        self.broadcast_parameters(model)
        ckpt = None if checkpoint_dir is None else (checkpoint_dir if self.world_size == 1 else os.path.join(checkpoint_dir, f"rank{self.rank}"))
        return Trainer(model, optimizer, loss_fn, ckpt, strategy=self, **trainer_kw)

    # ---- model-parallel GroupEmbedding for generic modules (gloo/nccl all-to-all; the fused engine uses parallel/p2p.py)
    def owner_of(self, table_index: int) -> int:
        return table_index % self.world_size

    def distributed_lookup(self, params, sp_ids, combiners, sp_weights=None) -> List[torch.Tensor]:
        """Model-parallel GroupEmbedding for generic modules over torch.distributed (gloo on CPU, nccl on GPU): table t lives on
        rank t % world; ONE id all-to-all, one lookup + combine per owned table, ONE row all-to-all back; the backward sends the
        output gradients to the owners, whose EmbeddingVariables record the sparse gradient (only the owner's optimizer ever updates
        a table).  See ``_GroupDistLookup``.  The flagship engine uses the fused NVLink kernels instead."""
        if sp_weights is None:
            sp_weights = [None] * len(params)
        if not hasattr(self, "_anchor"):
            self._anchor = torch.zeros(1, requires_grad=True)
        return list(_GroupDistLookup.apply(self._anchor, self, list(params), list(sp_ids), list(combiners), list(sp_weights)))

    def export_saved_model(self, *a, **k):
        from ..serving.export import export_saved_model
        return export_saved_model(*a, **k)
