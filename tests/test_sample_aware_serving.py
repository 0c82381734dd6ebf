"""Sample-aware graph compression as a pass over the exported op program (serving/export.py::compress_sample_aware): the user-side sub-graph of
a ranking request (user id + behaviour history -> pooled history, DSSM user tower, DIN history masking ...) runs ONCE per request at batch 1 and is
tiled where a per-candidate op consumes it.  Both native Processors must score a request exactly like the uncompressed export.

Reference: python/graph_optimizer/sample_awared_graph_compression.py:26 (user features sent once per request, tiled late)."""
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
from deeprec_b200.serving import Processor, compress_sample_aware, export_saved_model_program, taobao_user_columns

pytestmark = [pytest.mark.timeout(900)]


def _ranking_request(B, L, seed):
    """One user (id + history) x B candidate items: the user-side columns are identical in every row."""
    b = taobao_batch(B, L, 500, 3000, 40, seed=seed)
    for k in ("user", "hist_item", "hist_cat"):
        b[k] = b[k][:1].expand_as(b[k]).contiguous()
    return b


def _export_pair(tmp_path, name, L):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(11)
    model = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for sd in range(3):
        b = taobao_batch(64, L, 500, 3000, 40, seed=sd)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
    plain, comp = str(tmp_path / "plain"), str(tmp_path / "comp")
    export_saved_model_program(model, os.path.join(plain, "v1"), version=1, root=plain, max_len=L)
    export_saved_model_program(model, os.path.join(comp, "v1"), version=1, root=comp, max_len=L, sample_aware={"user_columns": taobao_user_columns(L)})
    return model, os.path.join(plain, "v1"), os.path.join(comp, "v1")


def test_pass_marks_the_user_side_subgraph_and_tiles_late():
    ops = [{"op": "slice", "out": "u", "in": ["emb"], "start": 0, "len": 16},
           {"op": "slice", "out": "q", "in": ["emb"], "start": 16, "len": 32},
           {"op": "linear", "out": "ut", "in": ["u"], "relu": True},
           {"op": "linear", "out": "it", "in": ["q"], "relu": True},
           {"op": "cosine", "out": "cos", "in": ["ut", "it"]}]
    new, out, n = compress_sample_aware(ops, "cos", user_columns=[0], emb_dim=16)
    kinds = [(o["op"], o["out"], bool(o.get("rows1"))) for o in new]
    assert kinds == [("slice", "u", True), ("slice", "q", False), ("linear", "ut", True), ("linear", "it", False), ("tile", "ut_tile", False),
                     ("cosine", "cos", False)]
    assert new[-1]["in"] == ["ut_tile", "it"] and out == "cos" and n == 2
    # nothing user-side -> the program is returned unchanged
    same, _, n0 = compress_sample_aware(ops, "cos", user_columns=[], emb_dim=16)
    assert n0 == 0 and [o["op"] for o in same] == [o["op"] for o in ops]


@pytest.mark.parametrize("name", ["dssm", "din", "mmoe"])
def test_compressed_program_scores_like_the_plain_one_on_the_cpu_processor(tmp_path, name):
    L, B = 12, 70
    model, plain_dir, comp_dir = _export_pair(tmp_path, name, L)
    meta = json.load(open(os.path.join(comp_dir, "saved_model.json")))
    assert meta["sample_aware"]["ops_at_batch_1"] >= 4 and any(o["op"] == "tile" for o in meta["program"])
    b = _ranking_request(B, L, seed=99)
    ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
    cfg = {"session_num": 1, "max_batch": 32, "model_update_interval_ms": 0}            # 70 rows -> chunks of 32: every chunk re-runs the user side
    plain, comp = Processor(plain_dir, cfg, device="cpu"), Processor(comp_dir, cfg, device="cpu")
    try:
        want, got = plain.predict(dense, ids), comp.predict(dense, ids)
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-5, np.abs(got - want).max()
        # only row 0 of a chunk's user-side columns is read by the compressed part: clients may pad the rest
        ids2 = ids.copy()
        ucols = taobao_user_columns(L)
        for start in range(0, B, 32):
            ids2[ucols, start + 1:min(B, start + 32)] = -1
        assert np.abs(comp.predict(dense, ids2) - want).max() < 1e-5
    finally:
        plain.close(); comp.close()


def test_compressed_dssm_and_din_on_the_emulated_gpu_processor(tmp_path):
    for name in ("dssm", "din"):
        L, B = 10, 36
        model, plain_dir, comp_dir = _export_pair(tmp_path / name, name, L)
        b = _ranking_request(B, L, seed=5)
        ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
        cfg = {"session_num": 1, "max_batch": 20, "model_update_interval_ms": 0}
        ref = Processor(plain_dir, cfg, device="cpu")
        emu = Processor(comp_dir, cfg, device="cuda_emu")
        try:
            want, got = ref.predict(dense, ids), emu.predict(dense, ids)
            assert np.isfinite(got).all() and np.abs(got - want).max() < 3e-2, (name, np.abs(got - want).max())
        finally:
            ref.close(); emu.close()


def test_compression_pays_on_the_cpu_processor(tmp_path):
    """DSSM at 256 candidates: the 4-layer user tower and the history pooling run once instead of 256 times."""
    L, B = 30, 256
    model, plain_dir, comp_dir = _export_pair(tmp_path, "dssm", L)
    b = _ranking_request(B, L, seed=1)
    ids = din_ids(b).numpy(); dense = np.zeros((B, 1), np.float32)
    cfg = {"session_num": 1, "max_batch": 256, "model_update_interval_ms": 0}
    t = {}
    for tag, d in (("plain", plain_dir), ("compressed", comp_dir)):
        p = Processor(d, cfg, device="cpu")
        try:
            for _ in range(5):
                p.predict(dense, ids)
            t0 = time.perf_counter()
            for _ in range(30):
                p.predict(dense, ids)
            t[tag] = (time.perf_counter() - t0) / 30
        finally:
            p.close()
    print(f"DSSM batch {B}: plain {t['plain'] * 1e3:.3f} ms, sample-aware {t['compressed'] * 1e3:.3f} ms")
    assert t["compressed"] < t["plain"] * 1.2          # 0.16 vs 0.275 ms on a quiet box; generous for a busy one
