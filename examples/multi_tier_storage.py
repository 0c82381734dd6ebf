"""DRAM + SSD two-tier EmbeddingVariable: a small DRAM tier caches the hot rows of a log-structured SSD store."""
import tempfile

import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr

with tempfile.TemporaryDirectory() as path:
    row_bytes = 4 * (8 + 8 + 4)                      # embedding + Adagrad accumulator + scalars
    ev = dr.get_embedding_variable("item_id", 8, ev_option=dr.EmbeddingVariableOption(storage_option=dr.StorageOption(
        dr.StorageType.DRAM_SSDHASH, storage_path=path, storage_size=[256 * row_bytes], cache_strategy=dr.CacheStrategy.LFU)))
    opt = dr.optim.AdagradOptimizer([], [ev], lr=0.1)
    g = torch.Generator().manual_seed(0)
    for step in range(20):
        ids = (torch.randn(128, generator=g).abs() * 400).long()       # skewed ids over ~1500 distinct values
        ev.lookup(ids).sum().backward(); opt.step()
    probe = torch.arange(0, 2000, 100)
    print("rows:", ev.total_count(), "tiers of probe ids (0 = DRAM, 1 = SSD, -1 = absent):", ev.lookup_tier(probe).tolist())
    print("tier statistics:", ev.table.tier_stats())
    assert ev.total_count() > 256 and (ev.lookup_tier(probe) == 1).any()
