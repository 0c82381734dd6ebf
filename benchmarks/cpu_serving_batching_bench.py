#!/usr/bin/env python
"""Request batching on the native CPU Processor, measured with the native load generator (benchmarks/cpu_serving_load.cc -- python client
threads saturate the GIL at ~10 k requests/s): N concurrent single-row callers, batching off vs on.  One JSON line per point."""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deeprec_b200 as dr  # noqa: E402
from deeprec_b200.data import criteo_batch  # noqa: E402
from deeprec_b200.models.zoo import build_model  # noqa: E402
from deeprec_b200.serving import export_saved_model_module  # noqa: E402

CARDS = [1000] * 26


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--requests", type=int, default=4000, help="per client thread")
    a = ap.parse_args()
    torch.manual_seed(0)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(3):
        d, ids, y = criteo_batch(2048, 13, CARDS, seed=s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    root = tempfile.mkdtemp()
    export_saved_model_module(model, root + "/v1", version=4)
    exe = os.path.join(root, "cpu_serving_load")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "benchmarks", "cpu_serving_load.cc"), "-o", exe, "-ldl"], check=True)
    lib = os.path.join(ROOT, "deeprec_b200", "lib", "libdeeprec_host.so")
    lines = []
    for threads in (8, 32):
        for name, cfg in (("off", {"session_num": 4}), ("max32/100us", {"session_num": 4, "enable_batching": True, "max_batch_size": 32, "batch_timeout_micros": 100}),
                          ("max32/100us strict", {"session_num": 4, "enable_batching": True, "max_batch_size": 32, "batch_timeout_micros": 100, "adaptive": False})):
            cfg = dict(cfg, model_update_interval_ms=0)
            r = subprocess.run([exe, lib, root + "/v1", json.dumps(cfg), str(threads), str(a.requests), "1", "13", "26"], capture_output=True, text=True)
            if r.returncode != 0:
                raise SystemExit(r.stderr)
            rec = json.loads(r.stdout.strip().splitlines()[-1])
            info = rec.pop("model_info")
            rec.update(metric="DLRM serving, native CPU Processor, concurrent single-row requests", batching=name, sessions=cfg["session_num"],
                       merged_batches=info["batching"]["merged_batches"], merged_requests=info["batching"]["merged_requests"], vcpus=os.cpu_count())
            print(json.dumps(rec), flush=True)
            lines.append(json.dumps(rec))
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
