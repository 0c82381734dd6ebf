"""Device placement optimisation of the GPU Processor (ModelConfig ``enable_device_placement_optimization``): the embedding tables stay in the HOST
engine (lookups on the caller's thread, one H2D copy of the rows per chunk), the dense part runs on the GPU -- tables larger than HBM, or one copy
of the tables per box instead of one per GPU.  Executed here on the CUDA-on-CPU emulation of the GPU runtime.

Reference: common_runtime/gpu/gpu_device_placement_pass.cc (embedding-layer ops on the CPU for GPU inference), config.proto:330,
ModelConfig ``enable_device_placement_optimization`` (docs/docs_en/Processor.md)."""
import os
import time

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch, taobao_batch
from deeprec_b200.models.rec_engine import din_ids
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_delta_program, export_saved_model_program

pytestmark = [pytest.mark.timeout(900)]
CARDS = [50, 1000, 7, 300] + [97] * 22
TOL = 3e-2


def _wait(pred, timeout=60.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.05)
    return False


def test_program_model_with_host_resident_tables(tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(4)
    model = build_model("deepfm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(4):
        d, ids, y = criteo_batch(96, 13, CARDS, seed=4 + s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=4, root=root)
    cfg = {"session_num": 2, "max_batch": 40, "checkpoint_dir": root, "model_update_interval_ms": 100, "enable_device_placement_optimization": True}
    host = Processor(os.path.join(root, "v1"), cfg, device="cuda_emu")
    dev = Processor(os.path.join(root, "v1"), {"session_num": 1, "max_batch": 40, "model_update_interval_ms": 0}, device="cuda_emu")
    try:
        assert host.model_info()["embedding_placement"] == "host" and dev.model_info()["embedding_placement"] == "device"
        model.eval()
        with torch.no_grad():
            ref = torch.sigmoid(model(d, ids)).numpy().copy()
        model.train()
        got = host.predict(d.numpy(), ids.numpy())                       # 96 rows -> chunks of 40
        assert np.abs(got - ref).max() < TOL
        assert np.abs(got - dev.predict(d.numpy(), ids.numpy())).max() < 1e-6      # same rows, same dense kernels: identical to device-resident tables
        ids2 = ids.clone(); ids2[:, :30] += 10 ** 9                       # unseen ids read the default rows of the host engine
        assert np.abs(host.predict(d.numpy()[:30], ids2.numpy()[:, :30]) - dev.predict(d.numpy()[:30], ids2.numpy()[:, :30])).max() < 1e-6
        for s in range(2):
            d2, i2, y2 = criteo_batch(96, 13, CARDS, seed=60 + s)
            loss = model.loss(d2, i2, y2); opt.zero_grad(); loss.backward(); opt.step()
        export_delta_program(model, root, base_version=4, version=6)      # copy-on-write import into the host tables
        assert _wait(lambda: host.model_info()["delta_version"] == 6)
        model.eval()
        with torch.no_grad():
            ref2 = torch.sigmoid(model(d, ids)).numpy().copy()
        assert np.abs(ref2 - ref).max() > 1e-4 and np.abs(host.predict(d.numpy(), ids.numpy()) - ref2).max() < TOL
    finally:
        host.close(); dev.close()


def test_sequence_program_with_host_resident_tables(tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(2)
    L, B = 10, 30
    model = build_model("din", device="cpu")
    b = taobao_batch(B, L, 500, 3000, 40, seed=1)
    b["hist_item"][:3] = -1; b["hist_cat"][:3] = -1
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L)
    model.eval()
    with torch.no_grad():
        ref = torch.sigmoid(model(b)).numpy().copy()
    ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
    p = Processor(os.path.join(root, "v1"), {"session_num": 1, "max_batch": 16, "model_update_interval_ms": 0, "embedding_placement": "host"}, device="cuda_emu")
    try:
        assert np.abs(p.predict(dense, ids) - ref).max() < TOL
    finally:
        p.close()


def test_dlrm_export_on_the_emulated_gpu_runtime_device_and_host_tables(tmp_path):
    """The flagship (non-program) serving path of serving_runtime.cu -- table kernels, folded MLP, dot interaction (SIMT variant on the
    emulation), head -- with device-resident and with host-resident tables (feature-major re-layout kernel), incl. a full hot swap."""
    from deeprec_b200.serving import export_saved_model_module
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(4):
        d, ids, y = criteo_batch(64, 13, CARDS, seed=s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    root = str(tmp_path)
    export_saved_model_module(model, os.path.join(root, "v1"), version=4, root=root)

    def ref_of():
        model.eval()
        with torch.no_grad():
            r = torch.sigmoid(model(d, ids)).numpy().copy()
        model.train()
        return r
    ref = ref_of()
    base = {"session_num": 1, "max_batch": 24, "checkpoint_dir": root, "model_update_interval_ms": 100}
    dev = Processor(os.path.join(root, "v1"), base, device="cuda_emu")
    host = Processor(os.path.join(root, "v1"), dict(base, enable_device_placement_optimization=True), device="cuda_emu")
    try:
        a, b2 = dev.predict(d.numpy(), ids.numpy()), host.predict(d.numpy(), ids.numpy())          # 64 rows -> chunks of 24
        assert np.abs(a - ref).max() < TOL and np.abs(b2 - a).max() < 1e-6
        for s in range(2):
            d2, i2, y2 = criteo_batch(64, 13, CARDS, seed=90 + s)
            loss = model.loss(d2, i2, y2); opt.zero_grad(); loss.backward(); opt.step()
        export_saved_model_module(model, os.path.join(root, "v2"), version=7, root=root)           # full update: fresh tables, warm-up, swap
        assert _wait(lambda: dev.model_info()["model_version"] == 7 and host.model_info()["model_version"] == 7)
        ref2 = ref_of()
        assert np.abs(ref2 - ref).max() > 1e-4
        assert np.abs(dev.predict(d.numpy(), ids.numpy()) - ref2).max() < TOL and np.abs(host.predict(d.numpy(), ids.numpy()) - ref2).max() < TOL
    finally:
        dev.close(); host.close()
