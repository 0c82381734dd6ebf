"""Train a model, serve it behind the HTTP front-end, call it with JSON, the compact wire format and protobuf."""
import numpy as np
import torch
from starlette.testclient import TestClient

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import SessionGroup
from deeprec_b200.serving.http_server import HttpClient, ServingBackend, create_app

torch.manual_seed(0)
model = build_model("deepfm", device="cpu")
opt = dr.optim.AdagradOptimizer(model, lr=0.05)
g = torch.Generator().manual_seed(1)
dense, ids, y = torch.randn(256, 13, generator=g), torch.randint(0, 500, (26, 256), generator=g), (torch.rand(256, generator=g) < 0.3).float()
for _ in range(5):
    opt.zero_grad(); model.loss(dense, ids, y).backward(); opt.step()

group = SessionGroup(model, session_num=4, select_session_policy="RR")
app = create_app({"deepfm": ServingBackend.from_session_group(group, version=1, extra_info={"num_dense": 13, "num_sparse": 26})})
# `deeprec_b200.serving.http_server.serve(...)` runs the same app under uvicorn; the test client keeps this example self-contained
with TestClient(app) as http:
    cli = HttpClient("http://testserver", "deepfm", session=http)
    a = cli.predict(dense[:8].numpy(), ids[:, :8].numpy())
    b = cli.predict_raw(dense[:8].numpy(), ids[:, :8].numpy())
    c = cli.predict_proto(dense[:8].numpy(), ids[:, :8].numpy(), per_feature=True)      # PredictRequest with inputs I1..I13, C1..C26
    print("json     :", np.round(a[:4], 4)); print("raw      :", np.round(b[:4], 4)); print("protobuf :", np.round(c[:4], 4))
    print(http.get("/v1/models/deepfm").json())
    assert np.allclose(a, b, atol=1e-6) and np.allclose(a, c, atol=1e-6)
