"""Synthetic Criteo-/Taobao-shaped batches (C++ generators in csrc/host/io_runtime.cc).

There are no datasets in the sandbox; ids follow a truncated power law so dedup / admission /
cache paths see realistic skew.  Layout: ids are feature-major ``[num_tables, batch]`` int64 (one
contiguous column per table -- the unit the model-parallel owner reads over NVLink)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence, Tuple

import torch

from .. import _native
from .._native import ptr


def criteo_batch(batch: int, num_dense: int, cardinalities: Sequence[int], seed: int = 0, alpha: float = 1.05,
                 threads: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    lib = _native.host()
    T = len(cardinalities)
    dense = torch.empty(batch, num_dense, dtype=torch.float32)
    ids = torch.empty(T, batch, dtype=torch.int64)
    labels = torch.empty(batch, dtype=torch.float32)
    cards = torch.tensor(list(cardinalities), dtype=torch.int64)
    nt = threads or min(8, os.cpu_count() or 1)
    lib.dr_gen_criteo(C.c_uint64(seed), batch, num_dense, T, ptr(cards), float(alpha), ptr(dense), ptr(ids), ptr(labels), nt)
    return dense, ids, labels


def taobao_batch(batch: int, max_len: int = 50, n_users: int = 1_000_000, n_items: int = 4_000_000, n_cats: int = 10_000,
                 seed: int = 0, alpha: float = 1.05):
    lib = _native.host()
    user = torch.empty(batch, dtype=torch.int64); item = torch.empty(batch, dtype=torch.int64); cat = torch.empty(batch, dtype=torch.int64)
    hi = torch.empty(batch, max_len, dtype=torch.int64); hc = torch.empty(batch, max_len, dtype=torch.int64)
    hl = torch.empty(batch, dtype=torch.int32); labels = torch.empty(batch, dtype=torch.float32)
    lib.dr_gen_taobao(C.c_uint64(seed), batch, max_len, n_users, n_items, n_cats, float(alpha), ptr(user), ptr(item), ptr(cat),
                      ptr(hi), ptr(hc), ptr(hl), ptr(labels))
    return dict(user=user, item=item, cat=cat, hist_item=hi, hist_cat=hc, hist_len=hl, labels=labels)
