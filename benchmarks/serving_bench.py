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
    ap.add_argument("--model", default="dlrm", help="dlrm (native architecture, trained on the fused engine) or an op-program model served by the same runtime: "
                    "wdl, deepfm, dcn, dcnv2, masknet (Criteo-shaped); din, dssm, esmm, mmoe, dbmtl, ple, simple_multitask (Taobao-shaped, 50-step history)")
    a = ap.parse_args()
    from deeprec_b200.data import criteo_batch
    from deeprec_b200.models.dlrm_engine import CRITEO_KAGGLE_CARDINALITIES, DLRMConfig, DLRMEngine
    from deeprec_b200.serving import Processor, ProcessorGroup, encode_request, export_saved_model
    cards = CRITEO_KAGGLE_CARDINALITIES
    root = tempfile.mkdtemp()
    if a.model == "dlrm":
        eng = DLRMEngine(DLRMConfig(batch_size=8192, cardinalities=cards))
        for s in range(8):
            d, ids, y = criteo_batch(8192, 13, cards, seed=s)
            eng.load_batch(d.cuda(), ids.cuda(), y.cuda()); eng.train_step()
        export_saved_model(eng, os.path.join(root, "v1"), version=8, root=root)
        del eng
        torch.cuda.empty_cache()

        def request(seed):
            d, ids, _ = criteo_batch(a.batch, 13, cards, seed=seed)
            return encode_request(d.numpy(), ids.numpy())
    else:                                   # op-program models: trained a few steps through the framework API on the host, exported, served on the GPU
        import deeprec_b200 as dr
        from deeprec_b200.models.zoo import TAOBAO_MODELS, build_model
        from deeprec_b200.serving import export_saved_model_program
        L = 50
        if a.model in TAOBAO_MODELS:
            from deeprec_b200.data import taobao_batch
            from deeprec_b200.models.rec_engine import din_ids
            model = build_model(a.model, device="cpu")
            opt = dr.optim.AdagradOptimizer(model, lr=0.05)
            for s in range(4):
                b = taobao_batch(2048, L, 100000, 400000, 1000, seed=s)
                loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
            export_saved_model_program(model, os.path.join(root, "v1"), version=8, root=root, max_len=L)

            def request(seed):
                b = taobao_batch(a.batch, L, 100000, 400000, 1000, seed=seed)
                return encode_request(np.zeros((a.batch, 1), np.float32), din_ids(b).numpy())
        else:
            small = [min(c, 200000) for c in cards]
            model = build_model(a.model, device="cpu", cardinalities=small)
            opt = dr.optim.AdagradOptimizer(model, lr=0.05)
            for s in range(4):
                d, ids, y = criteo_batch(2048, 13, small, seed=s)
                loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
            export_saved_model_program(model, os.path.join(root, "v1"), version=8, root=root)

            def request(seed):
                d, ids, _ = criteo_batch(a.batch, 13, small, seed=seed)
                return encode_request(d.numpy(), ids.numpy())
    cfg = {"session_num": a.sessions, "select_session_policy": a.policy, "max_batch": max(256, a.batch), "model_update_interval_ms": 0, "mlp_dtype": a.dtype}
    proc = Processor(os.path.join(root, "v1"), cfg, device="cuda") if a.gpus <= 1 else ProcessorGroup(os.path.join(root, "v1"), dict(cfg, gpu_ids_list=list(range(a.gpus))))
    reqs = [request(1000 + s) for s in range(16)]
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
    print(json.dumps({"metric": f"{a.model} serving (Processor C ABI, SessionGroup)", "model": a.model, "sessions": a.sessions, "client_threads": a.threads, "batch": a.batch,
                      "requests": int(n), "qps": n / wall, "samples_per_s": n * a.batch / wall, "p50_ms": float(allv[n // 2]),
                      "p99_ms": float(allv[min(n - 1, int(n * 0.99))]), "mean_ms": float(allv.mean()), "dtype": a.dtype, "n_gpus": a.gpus, "model_info": proc.model_info()}))
    proc.close()


if __name__ == "__main__":
    main()
