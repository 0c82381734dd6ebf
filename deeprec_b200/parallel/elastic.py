"""Elastic re-sharding of EmbeddingVariables when the number of workers changes at runtime.

Reference: contrib/elastic_grpc_server (PS scale-up/down: IsReadyScaling / ReadyToUpdate / UpdateServerDef /
FetchParamsMeta, elastic_training.proto:71-76) with the EV-side primitives ``GetShardedSnapshot`` / ``ExportAndRemove`` /
``RestoreFromKeysAndValues`` (embedding_var.h:521-599).  On a single NVSwitch node there is no PS tier; what remains is the
data movement: when the shard count goes N -> M every key whose owner changes is exported from the old owner (rows +
optimizer slots + freq/version), removed there, and imported by the new owner.  Works in-process (a list of shards, used by
tests and by single-process serving) and across ranks (torch.distributed object exchange).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch
import torch.distributed as dist

from ..embedding_variable import EmbeddingVariable


def default_owner(keys: torch.Tensor, num_shards: int) -> torch.Tensor:
    """Partitioned-EV rule (``key % 1000 % N``, the same rule the checkpoint N->M restore uses)."""
    return torch.remainder(torch.remainder(keys.to(torch.int64), 1000), num_shards)


def get_sharded_snapshot(ev: EmbeddingVariable, new_num_shards: int, owner_fn: Callable = default_owner) -> List[Dict[str, torch.Tensor]]:
    """``GetShardedSnapshot``: the table's admitted rows bucketed by their owner under the NEW shard count."""
    snap = ev.table.snapshot()
    own = owner_fn(snap["keys"], new_num_shards)
    out = []
    for s in range(new_num_shards):
        m = own == s
        out.append({"keys": snap["keys"][m], "rows": snap["rows"][m], "freqs": snap["freqs"][m], "versions": snap["versions"][m]})
    return out


def export_and_remove(ev: EmbeddingVariable, keep_shard: int, new_num_shards: int, owner_fn: Callable = default_owner) -> List[Dict[str, torch.Tensor]]:
    """``ExportAndRemove``: everything that no longer belongs to ``keep_shard`` leaves the table and is returned per new owner."""
    parts = get_sharded_snapshot(ev, new_num_shards, owner_fn)
    for s, p in enumerate(parts):
        if s != keep_shard and p["keys"].numel():
            ev.table.remove(p["keys"])
    return parts


def restore_from_keys_and_values(ev: EmbeddingVariable, part: Dict[str, torch.Tensor]) -> int:
    """``RestoreFromKeysAndValues``: import rows (full stride: embedding + optimizer slots) with their metadata."""
    if part["keys"].numel() == 0:
        return 0
    return ev.table.import_(part["keys"], part["rows"], part["freqs"], part["versions"])


def rescale_local(shards: Sequence[EmbeddingVariable], new_shards: Sequence[EmbeddingVariable], owner_fn: Callable = default_owner) -> int:
    """In-process N -> M: move every row to ``new_shards[owner(key, M)]`` (new_shards may reuse objects of ``shards``)."""
    M = len(new_shards)
    moved = 0
    exports = []
    for i, ev in enumerate(shards):
        keep = next((j for j, n in enumerate(new_shards) if n is ev), -1)
        exports.append((keep, export_and_remove(ev, keep, M, owner_fn)))
    for keep, parts in exports:
        for s, p in enumerate(parts):
            if s != keep:
                moved += restore_from_keys_and_values(new_shards[s], p)
    return moved


def rescale_distributed(ev: EmbeddingVariable, old_world: int, new_world: int, owner_fn: Callable = default_owner, group=None) -> int:
    """Across ranks (every rank of max(old, new) world calls this): rank r < old_world exports, rank s < new_world imports."""
    rank = dist.get_rank(group)
    W = dist.get_world_size(group)
    parts = export_and_remove(ev, rank if rank < new_world else -1, new_world, owner_fn) if rank < old_world else [None] * new_world
    gathered: List = [None] * W
    dist.all_gather_object(gathered, [None if p is None else {k: v.cpu() for k, v in p.items()} for p in parts], group=group)
    n = 0
    if rank < new_world:
        for src, plist in enumerate(gathered):
            if src != rank and plist[rank] is not None:
                n += restore_from_keys_and_values(ev, plist[rank])
    return n
