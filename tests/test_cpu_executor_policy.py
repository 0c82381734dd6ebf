"""Executor policies of the CPU Processor's op-program interpreter (csrc/host/cpu_serving.cc): NORMAL (program order, per-op teams), COST_MODEL
(per-op times traced over a window of requests, then the op DAG scheduled critical-path-first on the session's threads: independent towers /
experts side by side) and INLINE (caller's thread only) -- same numbers from all three.

Reference: ExecutorPolicy {NORMAL, COST_MODEL, INLINE} (core/protobuf/config.proto:19-26), common_runtime/executor.cc:414-478,1537,
costmodel*.{h,cc} / kernel_stat.h; env USE_INLINE_EXECUTOR, USE_COST_MODEL_EXECUTOR, START / STOP_NODE_STATS_STEP."""
import os
import threading

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import taobao_batch
from deeprec_b200.models.rec_engine import din_ids
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_saved_model_program

pytestmark = [pytest.mark.timeout(600)]


def _export(tmp_path, name, L):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(21)
    model = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for sd in range(3):
        b = taobao_batch(64, L, 500, 3000, 40, seed=sd)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
    d = str(tmp_path / name)
    export_saved_model_program(model, os.path.join(d, "v1"), version=1, root=d, max_len=L)
    return os.path.join(d, "v1")


@pytest.mark.parametrize("name", ["mmoe", "ple", "dssm", "din"])
def test_all_executor_policies_give_the_same_predictions(tmp_path, name):
    L = 12
    path = _export(tmp_path, name, L)
    b = taobao_batch(200, L, 500, 3000, 40, seed=77)
    ids = din_ids(b).numpy(); dense = np.zeros((200, 1), np.float32)
    base = {"session_num": 1, "max_batch": 256, "model_update_interval_ms": 0, "intra_op_parallelism_threads": 4}
    procs = {pol: Processor(path, dict(base, executor_policy=pol, start_node_stats_step=1, stop_node_stats_step=9), device="cpu") for pol in ("normal", "cost_model", "inline")}
    try:
        want_small, want_big = procs["normal"].predict(dense[:8], ids[:, :8]), procs["normal"].predict(dense, ids)
        cm = procs["cost_model"]
        for _ in range(12):                                          # warm-up requests + the tracing window
            assert np.abs(cm.predict(dense[:8], ids[:, :8]) - want_small).max() < 1e-6
        info = cm.model_info()["executor"]
        assert info["policy"] == "cost_model" and info["cost_model_ready"] and info["traced_runs"] >= 4 and info["ops"] > 5
        assert info["total_us"] >= info["critical_path_us"] > 0
        if name in ("mmoe", "ple"):
            assert info["dag_width"] >= 3                            # experts / towers are independent branches
        # scheduled small batches (several threads, any interleaving), per-op teams at large batch, 37 rows in between
        for n in (8, 1, 37):
            ref = procs["normal"].predict(dense[:n], ids[:, :n])
            for _ in range(5):
                assert np.abs(cm.predict(dense[:n], ids[:, :n]) - ref).max() < 1e-6
        assert np.abs(cm.predict(dense, ids) - want_big).max() < 1e-6
        assert np.abs(procs["inline"].predict(dense, ids) - want_big).max() < 1e-6
        assert procs["inline"].model_info()["executor"]["policy"] == "inline"
    finally:
        for p in procs.values():
            p.close()


def test_cost_model_executor_under_concurrent_sessions(tmp_path):
    """Two sessions schedule the same model's DAG at the same time (the plan is shared and read-only once ready; the run state is per session)."""
    L = 12
    path = _export(tmp_path, "mmoe", L)
    b = taobao_batch(16, L, 500, 3000, 40, seed=3)
    ids = din_ids(b).numpy(); dense = np.zeros((16, 1), np.float32)
    p = Processor(path, {"session_num": 2, "max_batch": 64, "model_update_interval_ms": 0, "intra_op_parallelism_threads": 3, "executor_policy": "cost_model",
                         "start_node_stats_step": 0, "stop_node_stats_step": 6}, device="cpu")
    ref = Processor(path, {"session_num": 1, "max_batch": 64, "model_update_interval_ms": 0}, device="cpu")
    try:
        want = ref.predict(dense, ids)
        for _ in range(10):
            p.predict(dense, ids)
        assert p.model_info()["executor"]["cost_model_ready"]
        errs = []

        def worker():
            for _ in range(40):
                if np.abs(p.predict(dense, ids) - want).max() > 1e-6:
                    errs.append(1)
        ts = [threading.Thread(target=worker) for _ in range(4)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs
    finally:
        p.close(); ref.close()
