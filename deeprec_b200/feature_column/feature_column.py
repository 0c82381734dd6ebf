"""Feature columns on EmbeddingVariables (python/feature_column/feature_column_v2.py in the reference).

``categorical_column_with_embedding`` (:2080), ``categorical_column_with_adaptive_embedding`` (:2089),
``categorical_column_with_multihash`` (:2103), ``categorical_column_with_hash_bucket`` / ``_identity``,
``sequence_categorical_column_with_embedding``, ``weighted_categorical_column``, ``embedding_column(..., do_fusion)`` (:658),
``shared_embedding_columns``, ``numeric_column``, ``group_embedding_column_scope`` (:4239), ``input_layer`` (:3189).

Columns are lightweight descriptors; ``InputLayer`` (an nn.Module) owns the tables and turns a feature dict into the
dense input tensor.  Columns created inside ``group_embedding_column_scope`` are looked up together through
``group_embedding_lookup_sparse`` (one fused launch on device / the model-parallel path under a CollectiveStrategy).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Union

import torch
from torch import nn

from ..config import EmbeddingVariableOption
from ..embedding_variable import EmbeddingVariable, get_embedding_variable, get_multihash_variable
from ..ops.embedding_ops import (SparseIds, adaptive_embedding_lookup_sparse, embedding_lookup, group_embedding_lookup_sparse,
                                 safe_embedding_lookup_sparse)

_GROUP_SCOPE: List[str] = []


@contextlib.contextmanager
def group_embedding_column_scope(name: str = "group"):
    """Embedding columns created inside the scope are fused into one GroupEmbedding lookup."""
    _GROUP_SCOPE.append(name)
    try:
        yield
    finally:
        _GROUP_SCOPE.pop()


@dataclass
class NumericColumn:
    key: str
    shape: Sequence[int] = (1,)
    normalizer_fn: Optional[object] = None

    @property
    def name(self):
        return self.key


@dataclass
class CategoricalColumn:
    key: str
    kind: str = "embedding"                    # embedding | hash_bucket | identity | adaptive | multihash
    num_buckets: int = 0
    dtype: torch.dtype = torch.int64
    ev_option: Optional[EmbeddingVariableOption] = None
    partition_num: Optional[int] = None
    multihash_dims: Optional[Sequence[Sequence[int]]] = None
    multihash_op: str = "add"
    is_sequence: bool = False

    @property
    def name(self):
        return self.key


@dataclass
class WeightedCategoricalColumn:
    categorical_column: CategoricalColumn
    weight_feature_key: str

    @property
    def name(self):
        return self.categorical_column.name + "_weighted_by_" + self.weight_feature_key

    @property
    def key(self):
        return self.categorical_column.key


@dataclass
class EmbeddingColumn:
    categorical_column: Union[CategoricalColumn, WeightedCategoricalColumn]
    dimension: int
    combiner: str = "mean"
    initializer: Optional[object] = None
    max_norm: Optional[float] = None
    trainable: bool = True
    do_fusion: bool = False
    shared_name: Optional[str] = None           # shared_embedding_columns
    group: Optional[str] = None

    @property
    def name(self):
        return self.categorical_column.name + "_embedding"


def numeric_column(key, shape=(1,), normalizer_fn=None):
    return NumericColumn(key, shape, normalizer_fn)


def categorical_column_with_embedding(key, dtype=torch.int64, partition_num=None, ev_option=None):
    return CategoricalColumn(key, "embedding", 0, dtype, ev_option, partition_num)


def categorical_column_with_hash_bucket(key, hash_bucket_size, dtype=torch.int64):
    return CategoricalColumn(key, "hash_bucket", hash_bucket_size, dtype)


def categorical_column_with_identity(key, num_buckets):
    return CategoricalColumn(key, "identity", num_buckets)


def categorical_column_with_adaptive_embedding(key, hash_bucket_size, dtype=torch.int64, partition_num=None, ev_option=None):
    return CategoricalColumn(key, "adaptive", hash_bucket_size, dtype, ev_option, partition_num)


def categorical_column_with_multihash(key, dims, complementary_strategy="Q-R", operation="add", dtype=torch.int64):
    return CategoricalColumn(key, "multihash", 0, dtype, None, None, dims, operation)


def sequence_categorical_column_with_embedding(key, dtype=torch.int64, partition_num=None, ev_option=None):
    return CategoricalColumn(key, "embedding", 0, dtype, ev_option, partition_num, is_sequence=True)


def sparse_column_with_embedding(column_name, dtype=torch.int64, partition_num=None, ev_option=None, **_ignored):
    """``tf.contrib.layers.sparse_column_with_embedding`` (contrib/layers/python/layers/feature_column.py): the contrib spelling of
    ``categorical_column_with_embedding`` -- an EmbeddingVariable-backed categorical column."""
    return categorical_column_with_embedding(column_name, dtype=dtype, partition_num=partition_num, ev_option=ev_option)


def weighted_categorical_column(categorical_column, weight_feature_key):
    return WeightedCategoricalColumn(categorical_column, weight_feature_key)


def embedding_column(categorical_column, dimension, combiner="mean", initializer=None, max_norm=None, trainable=True, do_fusion=False):
    return EmbeddingColumn(categorical_column, dimension, combiner, initializer, max_norm, trainable, do_fusion,
                           group=_GROUP_SCOPE[-1] if _GROUP_SCOPE else None)


def shared_embedding_columns(categorical_columns, dimension, combiner="mean", initializer=None, shared_embedding_collection_name=None,
                             max_norm=None, trainable=True):
    name = shared_embedding_collection_name or "_".join(sorted(c.name for c in categorical_columns)) + "_shared_embedding"
    return [EmbeddingColumn(c, dimension, combiner, initializer, max_norm, trainable, False, name, _GROUP_SCOPE[-1] if _GROUP_SCOPE else None)
            for c in categorical_columns]


def _to_sparse(x, dev=None) -> SparseIds:
    if isinstance(x, SparseIds):
        return x
    if not torch.is_tensor(x) and hasattr(x, "to_sparse"):      # data.DataFrameValue (ragged parquet column)
        return x.to_sparse()
    x = torch.as_tensor(x)
    if x.dim() == 1:
        return SparseIds.from_dense(x)
    return SparseIds.from_padded(x, pad_value=-1)


def _hash_bucket(ids: torch.Tensor, n: int) -> torch.Tensor:
    z = ids.to(torch.int64)
    z = (z ^ (z >> 30)) * -4658895280553007687          # splitmix-style mix, wraps in int64
    z = (z ^ (z >> 27)) * -7723592293110705685
    z = z ^ (z >> 31)
    return torch.remainder(z, n)


class InputLayer(nn.Module):
    """``tf.feature_column.input_layer``: owns the tables for the given columns and concatenates their outputs."""

    def __init__(self, feature_columns: Sequence[Union[NumericColumn, EmbeddingColumn]], device=None, name: str = "input_layer"):
        super().__init__()
        self.columns = list(feature_columns)
        self.device = device
        self.tables = nn.ModuleDict()
        self.hash_tables = nn.ModuleDict()
        self._key_of: Dict[int, str] = {}
        for c in self.columns:
            if isinstance(c, NumericColumn):
                continue
            cat = c.categorical_column.categorical_column if isinstance(c.categorical_column, WeightedCategoricalColumn) else c.categorical_column
            tname = (c.shared_name or f"{name}/{cat.key}_embedding").replace(".", "_")
            self._key_of[id(c)] = tname
            if tname in self.tables:
                continue
            if cat.kind in ("embedding", "adaptive"):
                self.tables[tname] = get_embedding_variable(tname, c.dimension, initializer=c.initializer, trainable=c.trainable,
                                                            partitioner=cat.partition_num, ev_option=cat.ev_option, device=device)
                if cat.kind == "adaptive":
                    self.hash_tables[tname] = nn.Embedding(cat.num_buckets, c.dimension, device=device)
            elif cat.kind == "multihash":
                self.tables[tname] = get_multihash_variable(tname, cat.multihash_dims, operation=cat.multihash_op, device=device)
            else:
                self.tables[tname] = nn.Embedding(cat.num_buckets, c.dimension, device=device)

    def embedding_variables(self) -> List[EmbeddingVariable]:
        return [m for m in self.modules() if isinstance(m, EmbeddingVariable)]

    def _lookup_one(self, c: EmbeddingColumn, features, adaptive_mask_tensors) -> torch.Tensor:
        wc = c.categorical_column if isinstance(c.categorical_column, WeightedCategoricalColumn) else None
        cat = wc.categorical_column if wc else c.categorical_column
        table = self.tables[self._key_of[id(c)]]
        sp = _to_sparse(features[cat.key])
        if wc is not None:
            w = features[wc.weight_feature_key]
            if isinstance(w, SparseIds):
                wv = w.values.to(torch.float32)
            else:
                w = torch.as_tensor(w, dtype=torch.float32)
                idr = features[cat.key]
                idr = torch.as_tensor(idr) if not isinstance(idr, SparseIds) else None
                wv = w[idr != -1] if (idr is not None and idr.dim() == 2 and w.shape == idr.shape) else w.reshape(-1)
            sp = SparseIds(sp.values, sp.row_ids, sp.batch_size, wv)
        if cat.kind == "hash_bucket":
            sp = SparseIds(_hash_bucket(sp.values, cat.num_buckets), sp.row_ids, sp.batch_size, sp.weights)
        if cat.is_sequence:
            raw = torch.as_tensor(features[cat.key])
            e = embedding_lookup(table, raw.clamp_min(0))
            return e * (raw >= 0).unsqueeze(-1).to(e.dtype)           # [B, T, D], padded steps zeroed
        if cat.kind == "adaptive":
            mask = None if adaptive_mask_tensors is None else adaptive_mask_tensors.get(cat.key)
            hs = SparseIds(_hash_bucket(sp.values, cat.num_buckets), sp.row_ids, sp.batch_size)
            return adaptive_embedding_lookup_sparse(self.hash_tables[self._key_of[id(c)]], table, sp, hs, sp.weights, c.combiner, c.max_norm, mask)
        return safe_embedding_lookup_sparse(table, sp, None, c.combiner, None, c.max_norm)

    def forward(self, features: Dict[str, object], adaptive_mask_tensors: Optional[Dict[str, torch.Tensor]] = None,
                cols_to_output_tensors: Optional[dict] = None) -> torch.Tensor:
        outs: Dict[int, torch.Tensor] = {}
        # ---- grouped columns: one fused lookup per group
        groups: Dict[str, List[EmbeddingColumn]] = {}
        for c in self.columns:
            if isinstance(c, EmbeddingColumn) and c.group is not None and not isinstance(c.categorical_column, WeightedCategoricalColumn) \
                    and c.categorical_column.kind == "embedding" and not c.categorical_column.is_sequence \
                    and isinstance(self.tables[self._key_of[id(c)]], EmbeddingVariable):
                groups.setdefault(c.group, []).append(c)
        for cols in groups.values():
            tabs = [self.tables[self._key_of[id(c)]] for c in cols]
            dev = tabs[0].device
            sps = [_to_sparse(features[c.categorical_column.key]).to(dev) for c in cols]
            res = group_embedding_lookup_sparse(tabs, sps, [c.combiner for c in cols])
            for c, r in zip(cols, res):
                outs[id(c)] = r
        parts = []
        for c in self.columns:
            if isinstance(c, NumericColumn):
                x = torch.as_tensor(features[c.key], dtype=torch.float32)
                x = x.view(x.shape[0], -1)
                if c.normalizer_fn is not None:
                    x = c.normalizer_fn(x)
                t = x
            else:
                t = outs[id(c)] if id(c) in outs else self._lookup_one(c, features, adaptive_mask_tensors)
            if cols_to_output_tensors is not None:
                cols_to_output_tensors[c.name] = t
            parts.append(t)
        dev = next((p.device for p in parts if p.device.type == "cuda"), parts[0].device)
        flat = [p.to(dev).flatten(1) for p in parts]
        return torch.cat(flat, dim=1)


def input_layer(features, feature_columns, adaptive_mask_tensors=None, cols_to_output_tensors=None, layer: Optional[InputLayer] = None):
    """Functional form; pass ``layer`` to reuse tables across calls (variables live in the layer)."""
    layer = layer or InputLayer(feature_columns)
    return layer(features, adaptive_mask_tensors, cols_to_output_tensors)
