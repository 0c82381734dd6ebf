"""HTTP serving front-end over a Python SessionGroup (CPU): JSON + raw wire format, model info, health, Prometheus metrics."""
import numpy as np
import torch

import deeprec_b200 as dr
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import SessionGroup
from deeprec_b200.serving.http_server import HttpClient, ServingBackend, create_app
from deeprec_b200.serving.processor import decode_response, encode_request


def test_http_predict_json_raw_and_metrics():
    from starlette.testclient import TestClient
    torch.manual_seed(0)
    model = build_model("wdl", device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    g = torch.Generator().manual_seed(1)
    dense = torch.randn(64, 13, generator=g); ids = torch.randint(0, 100, (26, 64), generator=g); y = (torch.rand(64, generator=g) < 0.4).float()
    for _ in range(3):
        opt.zero_grad(); model.loss(dense, ids, y).backward(); opt.step()
    group = SessionGroup(model, session_num=2)
    app = create_app({"wdl": ServingBackend.from_session_group(group, version=3)})
    ref = torch.sigmoid(group.run(dense, ids)).numpy()
    with TestClient(app) as http:
        assert http.get("/healthz").json() == {"status": "ok", "models": ["wdl"]}
        assert http.get("/v1/models/wdl").json()["model_version"] == 3
        r = http.post("/v1/models/wdl:predict", json={"dense": dense.tolist(), "ids": ids.tolist()})
        assert r.status_code == 200 and r.json()["model_version"] == 3
        assert np.allclose(np.asarray(r.json()["predictions"]), ref, atol=1e-6)
        r2 = http.post("/v1/models/wdl:predict", json={"dense": dense.tolist(), "ids": ids.t().tolist(), "ids_layout": "BT"})
        assert np.allclose(np.asarray(r2.json()["predictions"]), ref, atol=1e-6)
        raw = http.post("/v1/models/wdl:predict_raw", content=encode_request(dense.numpy(), ids.numpy()))
        probs, status, version = decode_response(raw.content)
        assert status == 200 and version == 3 and np.allclose(probs, ref, atol=1e-6)
        assert http.post("/v1/models/wdl:predict", json={"dense": [[1.0]], "ids": [[1, 2]]}).status_code == 400
        assert http.post("/v1/models/nope:predict", json={}).status_code == 404
        m = http.get("/metrics").text
        assert 'deeprec_requests_total{model="wdl",status="200"} 3.0' in m and 'status="400"} 1.0' in m
        assert "deeprec_request_latency_seconds_bucket" in m and "deeprec_request_batch_size_sum" in m
        # the python SDK speaks to the same endpoints (TestClient is requests-compatible)
        cli = HttpClient("http://testserver", "wdl", session=http)
        assert np.allclose(cli.predict(dense.numpy(), ids.numpy()), ref, atol=1e-6)
        assert np.allclose(cli.predict_raw(dense.numpy(), ids.numpy()), ref, atol=1e-6)


def test_any_zoo_model_round_trips_through_export_and_serves(tmp_path):
    """export_zoo_model / load_zoo_model: DIN (sequence features, three EmbeddingVariables) trained, exported, re-loaded in inference mode,
    served by a SessionGroup -- same predictions, and serving lookups of unseen ids create nothing."""
    from deeprec_b200.data import taobao_batch
    from deeprec_b200.serving import export_zoo_model, load_zoo_model
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model("din", device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for s in range(3):
        b = taobao_batch(128, 10, 1000, 2000, 50, seed=s)
        opt.zero_grad(); model.loss(b).backward(); opt.step()
    probe = taobao_batch(64, 10, 1000, 2000, 50, seed=99)
    model.eval()
    with torch.no_grad():
        ref = model(probe).clone()
    export_zoo_model(model, "din", str(tmp_path / "din_v3"), version=3)
    dr.embedding_variable.clear_registry()
    served, version = load_zoo_model(str(tmp_path / "din_v3"))
    assert version == 3
    group = SessionGroup(served, session_num=2)
    rows_before = served.item.total_count()
    got = group.run(probe)
    assert torch.allclose(got, ref, atol=1e-6)
    assert served.item.total_count() == rows_before == model.item.total_count()
