"""WorkQueue over parquet files -> ParquetDataset -> SmartStage (prefetch threads + packed host batch) -> training."""
import tempfile

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.data import ParquetDataset, SmartStageOptions, WorkQueue, smart_stage
from deeprec_b200.models.zoo import build_model

with tempfile.TemporaryDirectory() as d:
    rng = np.random.default_rng(0)
    files = []
    for part in range(4):
        cols = {"label": rng.integers(0, 2, 512).astype(np.float32)}
        cols.update({f"I{i}": rng.standard_normal(512).astype(np.float32) for i in range(1, 14)})
        cols.update({f"C{i}": rng.integers(0, 1000, 512).astype(np.int64) for i in range(1, 27)})
        files.append(f"{d}/part-{part}.parquet"); pq.write_table(pa.table(cols), files[-1])

    wq = WorkQueue(files, num_epochs=2, shuffle=True, seed=3)

    def read(path):                                     # one work item = one file
        for rec in ParquetDataset(path, batch_size=256):
            dense = torch.stack([rec[f"I{i}"] for i in range(1, 14)], 1)
            ids = torch.stack([rec[f"C{i}"] for i in range(1, 27)], 0)
            yield dense, ids, rec["label"]

    batches = smart_stage(wq.input_dataset(read), device=None, options=SmartStageOptions(capacity=4, num_threads=2))
    model = build_model("wdl", device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    n = 0
    for dense, ids, y in batches:
        loss = model.loss(dense, ids, y); opt.zero_grad(); loss.backward(); opt.step(); n += 1
    print("trained on", n, "batches; queue state", wq.state_dict())
    assert n == 2 * 4 * 2
