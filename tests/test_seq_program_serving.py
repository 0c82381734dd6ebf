"""DIEN (two GRUs around an attention-weight gate) and BST (one post-norm transformer encoder block) exported as op programs -- ``gru``,
``seq_last``, ``mha``, ``seq_mean``, per-position ``linear`` / ``layernorm``, ``din_attention`` in weights mode -- and served by the native CPU
Processor and by the GPU Processor's interpreter + kernels on the CUDA-on-CPU emulation: all 15 zoo models are now served natively.

Reference: the processor runs any SavedModel graph (serving/processor/serving/model_session.cc:377-386); models modelzoo/dien, modelzoo/bst."""
import json
import os
import time

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import taobao_batch
from deeprec_b200.models.rec_engine import din_ids
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_delta_program, export_saved_model_program, predict_pb

pytestmark = [pytest.mark.timeout(900)]


def _ref(model, b):
    model.eval()
    with torch.no_grad():
        p = torch.sigmoid(model(b)).numpy().copy()
    model.train()
    return p


def _wait(pred, timeout=60.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.05)
    return False


@pytest.mark.parametrize("name", ["dien", "bst"])
def test_recurrent_and_transformer_programs_on_both_native_processors(tmp_path, name):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(3)
    L, B = 12, 40
    model = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for sd in range(3):
        b = taobao_batch(B, L, 500, 3000, 40, seed=sd)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
    b["hist_item"][:4] = -1; b["hist_cat"][:4] = -1                     # four samples without any history
    b["hist_item"][4, 1:] = -1; b["hist_cat"][4, 1:] = -1               # one with a single valid position
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L)
    meta = json.load(open(os.path.join(root, "v1", "saved_model.json")))
    kinds = {o["op"] for o in meta["program"]}
    assert ({"gru", "seq_last", "din_attention"} if name == "dien" else {"mha", "seq_mean", "layernorm"}) <= kinds
    ref = _ref(model, b)
    ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
    cfg = {"session_num": 1, "max_batch": 25, "checkpoint_dir": root, "model_update_interval_ms": 100}         # 40 rows -> chunks of 25 + 15
    cpu = Processor(os.path.join(root, "v1"), cfg, device="cpu")
    emu = Processor(os.path.join(root, "v1"), dict(cfg, model_update_interval_ms=0), device="cuda_emu")
    try:
        host, got = cpu.predict(dense, ids), emu.predict(dense, ids)
        assert np.abs(host - ref).max() < 5e-5, np.abs(host - ref).max()                    # fp32 interpreter
        assert np.isfinite(got).all() and np.abs(got - ref).max() < 3e-2, np.abs(got - ref).max()     # bf16 activations
        assert np.abs(cpu.predict(dense[:1], ids[:, :1]) - ref[:1]).max() < 5e-5
        rc, pb = cpu.process(predict_pb.encode_predict_request(dense[:5], ids[:, :5]))
        assert rc == 200 and np.abs(predict_pb.decode_predict_response(pb)[0] - ref[:5]).max() < 5e-5
        # delta update: touched rows + every dense tensor of the program (GRU / attention / encoder weights included)
        for sd in range(2):
            t = taobao_batch(B, L, 500, 3000, 40, seed=40 + sd)
            loss = model.loss(t); opt.zero_grad(); loss.backward(); opt.step()
        export_delta_program(model, root, base_version=1, version=3, max_len=L)
        assert _wait(lambda: cpu.model_info()["delta_version"] == 3)
        ref2 = _ref(model, b)
        assert np.abs(ref2 - ref).max() > 1e-4 and np.abs(cpu.predict(dense, ids) - ref2).max() < 5e-5
    finally:
        cpu.close(); emu.close()


def test_cost_model_executor_runs_the_sequence_programs(tmp_path):
    """The DAG scheduler must respect the recurrences' dependencies (gru2 after the attention weights, the encoder's residual chain)."""
    dr.embedding_variable.clear_registry()
    torch.manual_seed(5)
    L = 10
    for name in ("dien", "bst"):
        model = build_model(name, device="cpu")
        root = str(tmp_path / name)
        export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L)
        b = taobao_batch(8, L, 500, 3000, 40, seed=2)
        ids = din_ids(b).numpy(); dense = np.zeros((8, 1), np.float32)
        ref = _ref(model, b)
        p = Processor(os.path.join(root, "v1"), {"session_num": 1, "max_batch": 16, "model_update_interval_ms": 0, "intra_op_parallelism_threads": 4,
                                                 "executor_policy": "cost_model", "start_node_stats_step": 0, "stop_node_stats_step": 4}, device="cpu")
        try:
            for _ in range(10):
                assert np.abs(p.predict(dense, ids) - ref).max() < 5e-5
            assert p.model_info()["executor"]["cost_model_ready"]
        finally:
            p.close()
