#!/usr/bin/env python
"""SessionGroup inference benchmark (BASELINE.json config #5): DLRM p50/p99 latency and QPS through the native Processor C ABI.

  python benchmarks/serving_bench.py --sessions 4 --threads 8 --batch 256 --requests 2000

Each client thread issues PredictRequests (batch rows of synthetic Criteo-shaped features) against ONE shared model; the
SessionGroup maps requests onto N sessions (stream + private buffers).  Reports per-request latency percentiles and
samples/s.  Prints one JSON line."""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, default=4)
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--requests", type=int, default=2000)
    ap.add_argument("--policy", default="RR")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp8"], help="MLP compute dtype of the serving runtime")
    ap.add_argument("--gpus", type=int, default=1, help="serve from this many GPUs (one Processor replica per GPU, ModelConfig gpu_ids_list)")
    a = ap.parse_args()
    from deeprec_b200.data import criteo_batch
    from deeprec_b200.models.dlrm_engine import CRITEO_KAGGLE_CARDINALITIES, DLRMConfig, DLRMEngine
    from deeprec_b200.serving import Processor, ProcessorGroup, encode_request, export_saved_model
    cards = CRITEO_KAGGLE_CARDINALITIES
    eng = DLRMEngine(DLRMConfig(batch_size=8192, cardinalities=cards))
    for s in range(8):
        d, ids, y = criteo_batch(8192, 13, cards, seed=s)
        eng.load_batch(d.cuda(), ids.cuda(), y.cuda()); eng.train_step()
    root = tempfile.mkdtemp()
    export_saved_model(eng, os.path.join(root, "v1"), version=8, root=root)
    del eng
    torch.cuda.empty_cache()
    cfg = {"session_num": a.sessions, "select_session_policy": a.policy, "max_batch": max(256, a.batch), "model_update_interval_ms": 0, "mlp_dtype": a.dtype}
    proc = Processor(os.path.join(root, "v1"), cfg) if a.gpus <= 1 else ProcessorGroup(os.path.join(root, "v1"), dict(cfg, gpu_ids_list=list(range(a.gpus))))
    reqs = []
    for s in range(16):
        d, ids, _ = criteo_batch(a.batch, 13, cards, seed=1000 + s)
        reqs.append(encode_request(d.numpy(), ids.numpy()))
    for r in reqs[:4]:
        assert proc.process(r)[0] == 200
    lat = [[] for _ in range(a.threads)]
    per = a.requests // a.threads

    def client(i):
        for k in range(per):
            t0 = time.perf_counter()
            rc, _ = proc.process(reqs[(i + k) % len(reqs)])
            lat[i].append((time.perf_counter() - t0) * 1e3)
            assert rc == 200
    ths = [threading.Thread(target=client, args=(i,)) for i in range(a.threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    allv = np.sort(np.concatenate([np.array(x) for x in lat]))
    n = allv.size
    print(json.dumps({"metric": "DLRM serving (Processor C ABI, SessionGroup)", "sessions": a.sessions, "client_threads": a.threads, "batch": a.batch,
                      "requests": int(n), "qps": n / wall, "samples_per_s": n * a.batch / wall, "p50_ms": float(allv[n // 2]),
                      "p99_ms": float(allv[min(n - 1, int(n * 0.99))]), "mean_ms": float(allv.mean()), "dtype": a.dtype, "n_gpus": a.gpus, "model_info": proc.model_info()}))
    proc.close()


if __name__ == "__main__":
    main()
