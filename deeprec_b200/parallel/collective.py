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
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import strategy as _strategy


class _DistLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, strategy, t, table, sp, combiner, weights):
        from ..ops.embedding_ops import SparseIds, embedding_lookup_sparse
        W, r, owner = strategy.world_size, strategy.rank, strategy.owner_of(t)
        objs = [None] * W
        w = weights if weights is not None else sp.weights
        dist.all_gather_object(objs, (sp.values.cpu(), sp.row_ids.cpu(), sp.batch_size, None if w is None else w.cpu()))
        outs = None
        if owner == r:
            with torch.enable_grad():
                outs = [embedding_lookup_sparse(table, SparseIds(v, rid, B, ww), None, combiner) for (v, rid, B, ww) in objs]
        recv = [None]
        dist.scatter_object_list(recv, [o.detach().cpu() for o in outs] if owner == r else None, src=owner)
        ctx.strategy, ctx.owner, ctx.outs = strategy, owner, outs
        return recv[0].to(sp.values.device if sp.values.is_cuda else "cpu")

    @staticmethod
    def backward(ctx, g):
        st = ctx.strategy
        gl = [None] * st.world_size if st.rank == ctx.owner else None
        dist.gather_object(g.detach().cpu(), gl, dst=ctx.owner)
        if st.rank == ctx.owner:
            torch.autograd.backward(ctx.outs, [x.to(o.device) for x, o in zip(gl, ctx.outs)])
        return (None,) * 7


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
        rank t % world; ids are all-gathered to the owner, the owner looks up + combines for every requester and scatters the
        results; the backward sends the output gradients back to the owner, whose EmbeddingVariable records the sparse gradient
        (so only the owner's optimizer ever updates the table).  The flagship engine uses the fused NVLink kernels instead."""
        if sp_weights is None:
            sp_weights = [None] * len(params)
        if not hasattr(self, "_anchor"):
            self._anchor = torch.zeros(1, requires_grad=True)
        return [_DistLookup.apply(self._anchor, self, t, p, s, c, w) for t, (p, s, c, w) in enumerate(zip(params, sp_ids, combiners, sp_weights))]

    def export_saved_model(self, *a, **k):
        from ..serving.export import export_saved_model
        return export_saved_model(*a, **k)
