"""Dynamic-dimension EmbeddingVariable (an id owns the first `blocknums[id]` blocks of its vector, the rest reads zero) and raw 64-bit
categorical ids without any hashing or vocabulary (the reference's `kaggle_categorical_int64` demo: Criteo hex strings -> int64 keys)."""
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr

dyn = dr.get_dynamic_dimension_embedding_variable("query", embedding_block_dimension=4, embedding_block_num=4)
ids = torch.tensor([11, 12, 13])
blocknums = torch.tensor([1, 2, 4])                      # e.g. from a frequency policy: rare ids get short vectors
e = dyn.lookup(ids, blocknums)
print("dynamic-dim rows (16 wide), non-zero widths:", [(row != 0).sum().item() for row in e])
assert [(row != 0).sum().item() for row in e] == [4, 8, 16]

ev = dr.get_embedding_variable("C14", embedding_dim=8)
keys = torch.tensor([int(h, 16) for h in ("68fd1e64", "80e26c9b", "fb936136", "7fffffffffffffff")], dtype=torch.int64)
opt = dr.optim.AdagradOptimizer([], [ev], lr=0.1)
out = ev.lookup(keys); opt.zero_grad(); out.sum().backward(); opt.step()
print("int64 keys stored as they are:", sorted(ev.export()[0].tolist()) == sorted(keys.tolist()), "| frequency", ev.get_frequency(keys).tolist())
