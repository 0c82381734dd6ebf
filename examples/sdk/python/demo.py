"""Python client demo: the same request over gRPC (protobuf) and over HTTP (JSON).

  python -m deeprec_b200.serving.serve --model ctr=/models/ctr/v7 --grpc_port 8501 --http_port 8500 &
  python examples/sdk/python/demo.py --grpc 127.0.0.1:8501 --http http://127.0.0.1:8500 --model ctr
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from deeprec_b200.serving.grpc_server import PredictClient  # noqa: E402
from deeprec_b200.serving.http_server import HttpClient  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grpc", default="127.0.0.1:8501")
    ap.add_argument("--http", default="http://127.0.0.1:8500")
    ap.add_argument("--model", default="ctr")
    ap.add_argument("--batch", type=int, default=4)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    dense = rng.standard_normal((a.batch, 13)).astype(np.float32)
    ids = rng.integers(0, 1000, (26, a.batch)).astype(np.int64)
    g = PredictClient(a.grpc, model=a.model)
    probs, version = g.predict(dense, ids)
    print("gRPC :", np.round(probs, 4), "model version", version, g.model_info().get("device"))
    h = HttpClient(a.http, a.model)
    print("HTTP :", np.round(h.predict(dense, ids), 4))
    g.close()


if __name__ == "__main__":
    main()
