"""Grouped lookup / update of host (DRAM) EmbeddingVariables: the CPU counterpart of the fused device GroupEmbedding path.

T tables of equal dim with one id per (table, sample) -- the Criteo layout -- are served by ONE native call that fills the
sample-major ``[B, T, D]`` tensor the interaction layers consume (no per-table tensors, no ``torch.stack``), parallel over
(table, key-chunk) tiles; the backward hands the whole ``[B, T, D]`` gradient to ONE native call that de-duplicates, sums and applies
every table (parallel over tables).  Semantics are those of T independent ``EmbeddingVariable.lookup`` + optimizer applies
(admission, frequency, version, default rows): the same C++ bodies run underneath.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from .. import _native
from .._native import ptr


def _handles(evs) -> "C.Array":
    return (C.c_void_p * len(evs))(*[ev.table.h for ev in evs])


class _HostGroupLookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, evs, ids):
        T, B = ids.shape
        out = torch.empty(B, T, evs[0].embedding_dim, dtype=torch.float32)
        _native.host().dr_host_group_lookup(_handles(evs), T, ptr(ids), B, ptr(out))
        ctx.evs, ctx.ids = evs, ids
        return out

    @staticmethod
    def backward(ctx, g):
        # the optimizer drains this list (DeepRecOptimizer.step): one grouped dedup + apply for all T tables
        ctx.evs[0]._group_pending.append((ctx.evs, ctx.ids, g.contiguous().float()))
        return None, None, None


def eligible(params: Sequence) -> bool:
    from ..embedding_variable import EmbeddingVariable, HostTable
    if not params or not all(isinstance(p, EmbeddingVariable) and p.device.type == "cpu" for p in params):
        return False
    d = params[0].embedding_dim
    return all(p.embedding_dim == d and isinstance(p.table, HostTable) for p in params)


def group_lookup_dense_host(params: Sequence, ids: torch.Tensor) -> Optional[torch.Tensor]:
    """ids [T, B] (feature-major, one id per table and sample) -> [B, T, D], or None when the tables are not plain host tables of one dim."""
    if ids.dim() != 2 or ids.shape[0] != len(params) or not eligible(params):
        return None
    evs: List = list(params)
    ids = ids.to(torch.int64).contiguous()
    train = torch.is_grad_enabled() and any(p.trainable and not p._inference for p in evs)
    if train:
        return _HostGroupLookup.apply(evs[0]._anchor, evs, ids)
    T, B = ids.shape
    out = torch.empty(B, T, evs[0].embedding_dim, dtype=torch.float32)
    _native.host().dr_host_group_lookup(_handles(evs), T, ptr(ids), B, ptr(out))
    return out


def apply_group_pending(evs: Sequence, hp) -> int:
    """Drain the grouped sparse gradients recorded on ``evs`` (called by the optimizer).  Returns the number of groups applied."""
    n = 0
    for ev in evs:
        pend = getattr(ev, "_group_pending", None)
        if not pend:
            continue
        for gevs, ids, g in pend:
            T, B = ids.shape
            _native.host().dr_host_group_apply_raw(_handles(gevs), T, ptr(ids), B, ptr(g), C.byref(hp))
            n += 1
        pend.clear()
    return n
