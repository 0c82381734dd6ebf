"""Train DLRM on CPU, export it, serve it with the native CPU Processor behind the gRPC PredictService, push a delta update to the live
model, and pin the two sessions to their own cores."""
import os
import tempfile
import time

import numpy as np
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import Processor, export_delta_module, export_saved_model_module
from deeprec_b200.serving.grpc_server import PredictClient, create_server

CARDS = [1000] * 26
root = tempfile.mkdtemp(prefix="deeprec_cpu_serving_")
torch.manual_seed(0)
model = build_model("dlrm", device="cpu", cardinalities=CARDS)
opt = dr.optim.AdagradOptimizer(model, lr=0.05)


def train(steps, seed):
    for s in range(steps):
        d, ids, y = criteo_batch(512, 13, CARDS, seed=seed + s)
        opt.zero_grad(); model.loss(d, ids, y).backward(); opt.step()


train(5, 0)
export_saved_model_module(model, os.path.join(root, "v5"), version=5, root=root)          # full version (BatchNorm folded into the GEMMs)

cpus = sorted(os.sched_getaffinity(0))
cfg = {"session_num": 2, "select_session_policy": "RR", "checkpoint_dir": root, "model_update_interval_ms": 100}
if len(cpus) >= 4:
    cfg["cpusets"] = f"{cpus[0]},{cpus[1]};{cpus[2]},{cpus[3]}"                            # SessionGroup.md: one core set per session
proc = Processor(os.path.join(root, "v5"), cfg, device="cpu")                              # libdeeprec_host.so: initialize / process / ...
server, port = create_server({"ctr": proc})
cli = PredictClient(f"127.0.0.1:{port}", model="ctr")

d, ids, _ = criteo_batch(64, 13, CARDS, seed=99)
p5, v = cli.predict(d.numpy(), ids.numpy())                                                # a real PredictRequest protobuf over gRPC
print("version", v, "probabilities", np.round(p5[:4], 4), cli.model_info()["cpusets"] or "(no cpusets)")

train(3, 100)
export_delta_module(model, root, base_version=5, version=8)                                  # only the rows touched since v5 (+ dense net)
t0 = time.time()
while cli.model_info()["delta_version"] != 8 and time.time() - t0 < 20:                     # the updater thread patches the LIVE tables
    time.sleep(0.05)
p8, v = cli.predict(d.numpy(), ids.numpy())
model.eval()
with torch.no_grad():
    ref = torch.sigmoid(model(d, ids)).numpy()
print("after the delta:", np.round(p8[:4], 4), "max |serving - module| =", float(np.abs(p8 - ref).max()))
assert np.abs(p8 - ref).max() < 1e-5 and np.abs(p8 - p5).max() > 0
cli.close(); server.stop(0); proc.close()
