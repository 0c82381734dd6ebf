#!/usr/bin/env python
"""Runs a few eager DLRM steps so `ncu -k regex:<kernel>` can capture the hot kernels (used by profiles/README.md recipe)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeprec_b200.data import criteo_batch  # noqa: E402
from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine  # noqa: E402

cfg = DLRMConfig(batch_size=65536, overlap_embedding=False)
eng = DLRMEngine(cfg)
for s in range(int(os.environ.get("STEPS", "3"))):
    d, ids, y = criteo_batch(cfg.batch_size, 13, cfg.cardinalities, seed=s)
    eng.load_batch(d.cuda(), ids.cuda(), y.cuda())
    eng.train_step()
torch.cuda.synchronize()
print("loss", eng.loss_value())
