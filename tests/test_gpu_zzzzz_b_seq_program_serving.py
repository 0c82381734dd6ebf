"""GPU Processor: DIEN and BST op programs (csrc/cuda/program_kernels.cu: k_prog_gru / k_prog_seq_last / k_prog_mha / k_prog_seq_mean, per-position
LINEAR / LAYERNORM on un-padded sequence buffers, din_attention in weights mode) against the modules and the CPU Processor.  Written in the last
session of round 2 without GPU access: the kernels and the interpreter were validated on the CUDA-on-CPU emulation (tests/test_seq_program_serving.py);
this file sorts LAST so that its first run on hardware cannot shadow anything else."""
import os

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import taobao_batch
from deeprec_b200.models.rec_engine import din_ids
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_saved_model_program

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("name", ["dien", "bst"])
def test_dien_and_bst_op_programs_on_the_gpu_processor(tmp_path, name):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(3)
    L, B = 20, 256
    model = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for sd in range(3):
        b = taobao_batch(B, L, 500, 3000, 40, seed=sd)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
    b["hist_item"][:5] = -1; b["hist_cat"][:5] = -1
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L)
    model.eval()
    with torch.no_grad():
        ref = torch.sigmoid(model(b)).numpy().copy()
    ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
    cfg = {"session_num": 2, "max_batch": 100, "model_update_interval_ms": 0}
    gpu = Processor(os.path.join(root, "v1"), cfg, device="cuda")
    cpu = Processor(os.path.join(root, "v1"), cfg, device="cpu")
    try:
        host, got = cpu.predict(dense, ids), gpu.predict(dense, ids)                     # 256 rows > max_batch: chunked
        assert np.abs(host - ref).max() < 5e-5
        assert np.isfinite(got).all() and np.abs(got - ref).max() < 3e-2, np.abs(got - ref).max()
        assert np.abs(gpu.predict(dense[:3], ids[:, :3]) - ref[:3]).max() < 3e-2
    finally:
        gpu.close(); cpu.close()
