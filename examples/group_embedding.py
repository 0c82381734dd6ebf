"""GroupEmbedding: N tables looked up and combined in one fused call; the feature-column scope does it implicitly."""
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.feature_column import categorical_column_with_embedding, embedding_column, group_embedding_column_scope, input_layer

evs = [dr.get_embedding_variable(f"C{i}", 8) for i in range(1, 5)]
sps = [dr.SparseIds.from_offsets(torch.tensor([1, 2, 3, 4, 5]), torch.tensor([0, 2, 2, 5])) for _ in evs]       # ragged: 2, 0, 3 ids per sample
outs = dr.group_embedding_lookup_sparse(evs, sps, ["sum", "mean", "sqrtn", "sum"])
print("group lookup:", [tuple(o.shape) for o in outs], "empty sample row is zero:", bool((outs[0][1] == 0).all()))

with group_embedding_column_scope("criteo"):
    cols = [embedding_column(categorical_column_with_embedding(f"F{i}"), 8, combiner="sum") for i in range(3)]
features = {f"F{i}": torch.randint(0, 1000, (16,)) for i in range(3)}
x = input_layer(features, cols)
print("input_layer with a group scope:", tuple(x.shape))
assert x.shape == (16, 24)
