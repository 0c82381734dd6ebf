#!/usr/bin/env python
"""Model-zoo training throughput on one GPU (BASELINE.json configs 3 and 4):

  --model deepfm   Criteo-shaped synthetic data, 26 EmbeddingVariables through ONE fused GroupEmbedding lookup, FM kernel + FusedMLP
  --model din      Taobao-shaped synthetic data; the user-behaviour (item) table draws ids from a 1B-row id space and lives in
                   multi-tier storage (--hbm_rows rows of HBM cache over host DRAM, optional SSD tier) with CounterFilter admission and
                   GlobalStep eviction enabled
Any other zoo model name works too.  Timing: CUDA events around K steps after W warm-up steps; the input H2D copy of every step
is inside the timed region.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deeprec_b200 as dr  # noqa: E402
from deeprec_b200.data import criteo_batch, taobao_batch  # noqa: E402
from deeprec_b200.models.zoo import CRITEO_MODELS, build_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="deepfm")
    ap.add_argument("--batch", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--optimizer", default="adagrad")
    ap.add_argument("--hbm_rows", type=int, default=1 << 20, help="DIN: HBM cache rows of the multi-tier item table")
    ap.add_argument("--id_space", type=int, default=1_000_000_000, help="DIN: size of the item id space")
    ap.add_argument("--ssd", action="store_true", help="DIN: add the SSD tier below DRAM (HBM_DRAM_SSDHASH)")
    ap.add_argument("--filter_freq", type=int, default=2)
    a = ap.parse_args()
    dev = torch.device("cuda")
    name = a.model.lower()
    criteo = name in CRITEO_MODELS or name == "dlrm"
    if criteo:
        opt_ev = dr.EmbeddingVariableOption(storage_option=dr.StorageOption(dr.StorageType.HBM))
        model = build_model(name, ev_option=opt_ev, device=dev, group_embedding=True)
        cards = [1_000_000] * 26
        gen = lambda s: criteo_batch(a.batch, 13, cards, seed=s)
    else:
        row_bytes = 4 * 16 * 2
        st = dr.StorageType.HBM_DRAM_SSDHASH if a.ssd else dr.StorageType.HBM_DRAM
        opt_ev = dr.EmbeddingVariableOption(storage_option=dr.StorageOption(dr.StorageType.HBM))
        model = build_model(name, ev_option=opt_ev, device=dev)
        # re-create the behaviour table as a multi-tier, filtered, evicting EmbeddingVariable
        big = dr.EmbeddingVariableOption(filter_option=dr.CounterFilter(a.filter_freq), evict_option=dr.GlobalStepEvict(100000),
                                         storage_option=dr.StorageOption(st, storage_size=(a.hbm_rows * row_bytes, 8 << 30),
                                                                         cache_strategy=dr.CacheStrategy.LFU))
        model.item = dr.get_embedding_variable(f"{name}/item_multitier", 16, ev_option=big, device=dev)
        gen = lambda s: taobao_batch(a.batch, 50, 10_000_000, a.id_space, 10_000, seed=s)
    opt = dr.optim.make_optimizer(a.optimizer, model, lr=0.01)

    def to_dev(b):
        if isinstance(b, dict):
            return {k: v.pin_memory().to(dev, non_blocking=True) for k, v in b.items()}
        return tuple(t.pin_memory().to(dev, non_blocking=True) for t in b)

    def step(b):
        b = to_dev(b)
        loss = model.loss(b) if isinstance(b, dict) else model.loss(*b)
        opt.zero_grad(); loss.backward(); opt.step()
        return loss

    batches = [gen(s) for s in range(a.warmup + a.steps)]
    for i in range(a.warmup):
        step(batches[i])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        loss = step(batches[a.warmup + i])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    out = {"metric": f"{name} training samples/s (1 GPU, framework API, H2D inside the timed region)", "value": a.batch / ms * 1e3, "unit": "samples/s",
           "ms_per_step": ms, "wall_ms_per_step": (time.perf_counter() - t0) * 1e3 / a.steps, "batch": a.batch, "steps": a.steps, "warmup": a.warmup,
           "final_loss": float(loss.item()), "dtype": "bf16 GEMMs / fp32 master weights", "data": "synthetic"}
    if not criteo:
        t = model.item.table
        out["multi_tier"] = {k: (float(v) if not isinstance(v, dict) else v) for k, v in t.cache_stats().items()}
        out["item_table"] = {"admitted_rows": int(model.item.total_count()), "id_space": a.id_space, "hbm_rows": a.hbm_rows}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
