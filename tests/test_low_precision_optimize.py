"""tools/low_precision_optimize: ``optimize`` (BF16 / FP16 / INT8 of an exported saved model, embeddings included, per-node ``opt_config``, calibration
set) and ``convert_ckpt`` (the same decisions on a later delta); both native Processors load and hot-update the optimised model.

Reference: tools/low_precision_optimize/{low_precision_optimize.py, README.md} (optimize / convert_ckpt, 'Optimization Result' log)."""
import json
import os
import time

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_delta_program, export_saved_model_program
from deeprec_b200.tools.low_precision_optimize import convert_ckpt, optimize

CARDS = [50, 1000, 7, 300] + [97] * 22
pytestmark = [pytest.mark.timeout(600)]


def _size(d):
    return sum(os.path.getsize(os.path.join(r, f)) for r, _, fs in os.walk(d) for f in fs)


def _wait(pred, timeout=30.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.05)
    return False


def test_optimize_saved_model_and_deltas(tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(4)
    model = build_model("deepfm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(4):
        d, ids, y = criteo_batch(256, 13, CARDS, seed=4 + s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    root = str(tmp_path / "fp32")
    export_saved_model_program(model, os.path.join(root, "v1"), version=4, root=root)
    calib = [(d.numpy()[:64], ids.numpy()[:, :64]), (d.numpy()[64:128], ids.numpy()[:, 64:128])]
    logs = []
    out = {}
    for dt, tol in (("BF16", 5e-3), ("FP16", 2e-3), ("INT8", 2e-2)):
        out[dt] = optimize(os.path.join(root, "v1"), str(tmp_path / dt / "v1"), data_type=dt, calib_data=calib, max_drift=tol, log=logs.append)
        assert out[dt]["max_probability_drift"] < tol
    assert logs[0] == "Optimization Result:" and any(l.startswith("Optimize embedding to BF16: table/0-values") for l in logs)
    assert any(l.startswith("Optimize dense to INT8: prog/") for l in logs)
    assert out["BF16"]["ratio"] < 0.56 and out["INT8"]["ratio"] < 0.36                       # the files really shrink (keys / defaults stay as they are)
    assert _size(str(tmp_path / "INT8")) < 0.45 * _size(os.path.join(root, "v1"))
    meta = json.load(open(tmp_path / "INT8" / "v1" / "saved_model.json"))
    assert meta["low_precision"]["data_type"] == "INT8" and meta["low_precision"]["plan"]["table/3"] == "INT8"
    # per-node configuration: only what is named is touched; unknown names are an error
    part = optimize(os.path.join(root, "v1"), str(tmp_path / "part" / "v1"), opt_config={"table/1": "INT8", "table/2": "FP16"}, log=None)
    assert sorted(part["plan"].items()) == [("table/1-default", "int8"), ("table/1-values", "int8"), ("table/2-default", "fp16"), ("table/2-values", "fp16")]
    with pytest.raises(KeyError):
        optimize(os.path.join(root, "v1"), str(tmp_path / "bad" / "v1"), opt_config={"no/such/node": "BF16"}, log=None)
    # the optimised model serves on both runtimes and takes a converted delta
    model.eval()
    with torch.no_grad():
        ref = torch.sigmoid(model(d, ids)).numpy().copy()
    model.train()
    iroot = str(tmp_path / "INT8")
    cpu = Processor(os.path.join(iroot, "v1"), {"session_num": 1, "max_batch": 256, "checkpoint_dir": iroot, "model_update_interval_ms": 100}, device="cpu")
    emu = Processor(os.path.join(iroot, "v1"), {"session_num": 1, "max_batch": 64, "model_update_interval_ms": 0}, device="cuda_emu")
    try:
        assert np.abs(cpu.predict(d.numpy(), ids.numpy()) - ref).max() < 2e-2
        assert np.abs(emu.predict(d.numpy()[:64], ids.numpy()[:, :64]) - ref[:64]).max() < 4e-2
        for s in range(2):
            d2, i2, y2 = criteo_batch(256, 13, CARDS, seed=70 + s)
            loss = model.loss(d2, i2, y2); opt.zero_grad(); loss.backward(); opt.step()
        export_delta_program(model, root, base_version=4, version=6)                         # fp32 delta of the training side ...
        delta = os.path.join(root, ".incr", "delta-6")
        os.makedirs(os.path.join(iroot, ".incr"), exist_ok=True)
        r = convert_ckpt(delta, os.path.join(iroot, ".incr", "delta-6"), os.path.join(iroot, "v1"))     # ... shipped in the optimised format
        assert r["ratio"] < 0.5
        vers = json.load(open(os.path.join(root, "serving_versions.json")))
        for e in vers.get("deltas", []):
            e["prefix"] = e["prefix"].replace(root, iroot)
        vers["full"]["dir"] = os.path.join(iroot, "v1")
        json.dump(vers, open(os.path.join(iroot, "serving_versions.json"), "w"))
        assert _wait(lambda: cpu.model_info()["delta_version"] == 6)
        model.eval()
        with torch.no_grad():
            ref2 = torch.sigmoid(model(d, ids)).numpy().copy()
        assert np.abs(ref2 - ref).max() > 1e-4 and np.abs(cpu.predict(d.numpy(), ids.numpy()) - ref2).max() < 2e-2
    finally:
        cpu.close(); emu.close()
