"""Any zoo model -> op program -> native Processor.  MMoE with the cost-model executor (independent experts / towers side by side for
latency-bound requests), DSSM with sample-aware graph compression (the user tower once per ranking request), BST (transformer block)."""
import os
import tempfile
import time

import numpy as np
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.data import taobao_batch
from deeprec_b200.models.rec_engine import din_ids
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_saved_model_program, taobao_user_columns

L = 20


def export(name, **kw):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model(name, device="cpu")
    root = tempfile.mkdtemp()
    export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L, **kw)
    return model, os.path.join(root, "v1")


def p50(proc, dense, ids, n=200):
    for _ in range(50):
        proc.predict(dense, ids)
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); proc.predict(dense, ids); ts.append(time.perf_counter() - t0)
    return sorted(ts)[n // 2] * 1e6


b = taobao_batch(64, L, 500, 3000, 40, seed=1)
ids, dense = din_ids(b).numpy(), np.zeros((64, 1), np.float32)

# ---- executor policies
model, path = export("mmoe")
cfg = {"session_num": 1, "max_batch": 64, "model_update_interval_ms": 0, "intra_op_parallelism_threads": 4}
for policy in ("normal", "cost_model"):
    p = Processor(path, dict(cfg, executor_policy=policy, start_node_stats_step=2, stop_node_stats_step=20), device="cpu")
    us = p50(p, dense[:4], np.ascontiguousarray(ids[:, :4]))
    print(f"MMoE, 4 rows, executor_policy={policy:10s}: p50 {us:6.1f} us", {k: v for k, v in p.model_info()["executor"].items() if k in ("dag_width", "team", "parallel")})
    p.close()

# ---- sample-aware compression: one user, 64 candidate items
for k in ("user", "hist_item", "hist_cat"):
    b[k] = b[k][:1].expand_as(b[k]).contiguous()
ids = din_ids(b).numpy()
for tag, kw in (("plain", {}), ("sample-aware", {"sample_aware": {"user_columns": taobao_user_columns(L)}})):
    model, path = export("dssm", **kw)
    p = Processor(path, cfg, device="cpu")
    print(f"DSSM, 64 candidates, {tag:12s}: p50 {p50(p, dense, ids):6.1f} us")
    p.close()

# ---- a transformer block as an op program
model, path = export("bst")
p = Processor(path, cfg, device="cpu")
model.eval()
with torch.no_grad():
    ref = torch.sigmoid(model(b)).numpy()
print("BST native vs module, max |diff| =", float(np.abs(p.predict(dense, ids) - ref).max()))
p.close()
