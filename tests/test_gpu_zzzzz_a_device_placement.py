"""GPU Processor with host-resident tables (``enable_device_placement_optimization``): rows looked up by the host engine, one H2D copy per chunk,
dense part on the GPU.  Written without GPU access in the last session of round 2 (validated on the CUDA-on-CPU emulation:
tests/test_device_placement_serving.py); sorts last."""
import os

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_saved_model_module, export_saved_model_program

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
CARDS = [50, 1000, 7, 300] + [97] * 22


@pytest.mark.parametrize("name", ["dlrm", "deepfm"])
def test_host_resident_tables_match_device_resident_tables(tmp_path, name):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model(name, device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(4):
        d, ids, y = criteo_batch(512, 13, CARDS, seed=s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    root = str(tmp_path)
    (export_saved_model_module if name == "dlrm" else export_saved_model_program)(model, os.path.join(root, "v1"), version=1, root=root)
    model.eval()
    with torch.no_grad():
        ref = torch.sigmoid(model(d, ids)).numpy().copy()
    cfg = {"session_num": 2, "max_batch": 200, "model_update_interval_ms": 0}
    dev = Processor(os.path.join(root, "v1"), cfg, device="cuda")
    host = Processor(os.path.join(root, "v1"), dict(cfg, enable_device_placement_optimization=True), device="cuda")
    try:
        a, b = dev.predict(d.numpy(), ids.numpy()), host.predict(d.numpy(), ids.numpy())
        assert host.model_info()["embedding_placement"] == "host"
        assert np.abs(a - ref).max() < 3e-2 and np.abs(b - a).max() < 1e-6
    finally:
        dev.close(); host.close()
