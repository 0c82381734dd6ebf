"""HTTP front-end for the serving runtimes (the reference ships its processor behind EAS / TF-Serving style endpoints and has
client SDK demos in serving/sdk/; this is the in-repo equivalent): FastAPI app with

  POST /v1/models/{name}:predict     JSON ``{"dense": [[...]], "ids": [[...]]}``  (ids feature-major [T][B], or sample-major with
                                     ``"ids_layout": "BT"``) -> ``{"predictions": [...], "model_version": v}``
  POST /v1/models/{name}:predict_proto protobuf ``PredictRequest`` -> ``PredictResponse`` (wire format of the reference's predict.proto)
  POST /v1/models/{name}:predict_raw the binary wire format of ``process()`` (``encode_request`` / ``decode_response``) as-is
  GET  /v1/models/{name}             model / version / session information (``get_serving_model_info``)
  GET  /healthz                      liveness
  GET  /metrics                      Prometheus text format: request / failure counters, latency histogram, batch-size histogram

Backends: the native ``Processor`` (C ABI, GPU) or any callable ``predict(dense: np.ndarray, ids: np.ndarray) -> np.ndarray`` such as a
Python ``SessionGroup`` around a zoo model (CPU or GPU).
"""

import time
from typing import Callable, Dict, Optional

import numpy as np

from .processor import decode_response, encode_request


class ServingBackend:
    """Uniform view of a model for the HTTP layer."""

    def __init__(self, predict: Callable[[np.ndarray, np.ndarray], np.ndarray], info: Optional[Callable[[], dict]] = None,
                 process_raw: Optional[Callable[[bytes], tuple]] = None):
        self.predict, self.info, self.process_raw = predict, info or (lambda: {}), process_raw

    @classmethod
    def from_processor(cls, proc) -> "ServingBackend":
        return cls(proc.predict, proc.model_info, proc.process)

    @classmethod
    def from_session_group(cls, group, to_inputs: Optional[Callable] = None, version: int = 0, extra_info: Optional[dict] = None) -> "ServingBackend":
        import torch

        def predict(dense: np.ndarray, ids: np.ndarray) -> np.ndarray:
            d, i = torch.from_numpy(np.ascontiguousarray(dense, dtype=np.float32)), torch.from_numpy(np.ascontiguousarray(ids, dtype=np.int64))
            args = to_inputs(d, i) if to_inputs else (d, i)
            out = group.run(*args)
            return torch.sigmoid(out).reshape(-1).cpu().numpy()
        return cls(predict, lambda: {"model_version": version, "sessions": len(group.sessions), **(extra_info or {})})


def create_app(backends: Dict[str, ServingBackend]):
    from fastapi import FastAPI, HTTPException, Request, Response
    from prometheus_client import CollectorRegistry, Counter, Histogram, generate_latest

    app = FastAPI(title="deeprec_b200 serving")
    reg = CollectorRegistry()
    req_total = Counter("deeprec_requests_total", "prediction requests", ["model", "status"], registry=reg)
    latency = Histogram("deeprec_request_latency_seconds", "end-to-end request latency", ["model"], registry=reg,
                        buckets=(1e-4, 2.5e-4, 5e-4, 1e-3, 2.5e-3, 5e-3, 1e-2, 2.5e-2, 5e-2, 0.1, 0.25, 1.0))
    batch = Histogram("deeprec_request_batch_size", "samples per request", ["model"], registry=reg, buckets=(1, 8, 32, 128, 512, 2048, 8192))
    app.state.registry = reg

    def _backend(name: str) -> ServingBackend:
        if name not in backends:
            raise HTTPException(status_code=404, detail=f"unknown model {name!r}; serving: {sorted(backends)}")
        return backends[name]

    @app.get("/healthz")
    def healthz():
        return {"status": "ok", "models": sorted(backends)}

    @app.get("/metrics")
    def metrics():
        return Response(generate_latest(reg), media_type="text/plain; version=0.0.4")

    @app.get("/v1/models/{name}")
    def model_info(name: str):
        return _backend(name).info()

    @app.post("/v1/models/{name}:predict")
    async def predict(name: str, request: Request):
        be = _backend(name)
        t0 = time.perf_counter()
        try:
            body = await request.json()
            dense = np.asarray(body["dense"], dtype=np.float32)
            ids = np.asarray(body["ids"], dtype=np.int64)
            if body.get("ids_layout", "TB").upper() == "BT":
                ids = np.ascontiguousarray(ids.T)
            if dense.ndim != 2 or ids.ndim != 2 or ids.shape[1] != dense.shape[0]:
                raise ValueError(f"expected dense [B, num_dense] and ids [T, B]; got {dense.shape} and {ids.shape}")
            probs = be.predict(dense, ids)
        except HTTPException:
            raise
        except (KeyError, ValueError, TypeError) as e:
            req_total.labels(name, "400").inc()
            raise HTTPException(status_code=400, detail=str(e))
        except Exception as e:          # backend failure
            req_total.labels(name, "500").inc()
            raise HTTPException(status_code=500, detail=str(e))
        req_total.labels(name, "200").inc()
        latency.labels(name).observe(time.perf_counter() - t0)
        batch.labels(name).observe(dense.shape[0])
        return {"predictions": [float(p) for p in probs], "model_version": be.info().get("model_version", 0)}

    @app.post("/v1/models/{name}:predict_raw")
    async def predict_raw(name: str, request: Request):
        be = _backend(name)
        payload = await request.body()
        t0 = time.perf_counter()
        if be.process_raw is not None:
            rc, out = be.process_raw(payload)
        else:
            rc, out = _emulate_wire(be, payload)
        req_total.labels(name, str(rc)).inc()
        if rc == 200:
            latency.labels(name).observe(time.perf_counter() - t0)
        return Response(out, status_code=rc, media_type="application/octet-stream")

    @app.post("/v1/models/{name}:predict_proto")
    async def predict_proto(name: str, request: Request):
        from . import predict_pb
        be = _backend(name)
        payload = await request.body()
        t0 = time.perf_counter()
        try:
            if be.process_raw is not None:      # the native runtime parses protobuf itself
                rc, out = be.process_raw(payload)
            else:
                info = be.info()
                wire = predict_pb.request_to_wire(payload, int(info["num_dense"]), int(info["num_sparse"]))
                rc, out = _emulate_wire(be, wire)
                if rc == 200:
                    out = predict_pb.response_from_wire(out, payload)
        except (KeyError, ValueError) as e:
            req_total.labels(name, "400").inc()
            raise HTTPException(status_code=400, detail=str(e))
        req_total.labels(name, str(rc)).inc()
        if rc == 200:
            latency.labels(name).observe(time.perf_counter() - t0)
        return Response(out, status_code=rc, media_type="application/x-protobuf")

    return app


def _emulate_wire(be: "ServingBackend", payload: bytes):
    """The compact wire format on top of a ``predict()`` callable (backends without a native ``process``)."""
    import struct
    try:
        magic, ver, b, nd, ns, _ = struct.unpack_from("<6I", payload, 0)
        dense = np.frombuffer(payload, np.float32, b * nd, 24).reshape(b, nd)
        ids = np.frombuffer(payload, np.int64, ns * b, 24 + 4 * b * nd).reshape(ns, b)
        probs = np.asarray(be.predict(dense, ids), dtype=np.float32)
        return 200, struct.pack("<4Iq", 0x53525244, b, 200, 0, int(be.info().get("model_version", 0))) + probs.tobytes()
    except Exception:
        return 500, b""


def serve(backends: Dict[str, ServingBackend], host: str = "127.0.0.1", port: int = 8500, **uvicorn_kw) -> None:
    import uvicorn
    uvicorn.run(create_app(backends), host=host, port=port, **uvicorn_kw)


class HttpClient:
    """Minimal python SDK (serving/sdk/python in the reference): JSON and raw calls."""

    def __init__(self, base_url: str, model: str, session=None):
        import requests
        self.url, self.model, self.http = base_url.rstrip("/"), model, session or requests.Session()

    def predict(self, dense: np.ndarray, ids: np.ndarray) -> np.ndarray:
        r = self.http.post(f"{self.url}/v1/models/{self.model}:predict", json={"dense": np.asarray(dense).tolist(), "ids": np.asarray(ids).tolist()})
        r.raise_for_status()
        return np.asarray(r.json()["predictions"], dtype=np.float32)

    def predict_proto(self, dense: np.ndarray, ids: np.ndarray, per_feature: bool = False) -> np.ndarray:
        from .predict_pb import decode_predict_response, encode_predict_request
        r = self.http.post(f"{self.url}/v1/models/{self.model}:predict_proto", data=encode_predict_request(dense, ids, per_feature=per_feature))
        r.raise_for_status()
        return decode_predict_response(r.content)[0]

    def predict_raw(self, dense: np.ndarray, ids: np.ndarray) -> np.ndarray:
        r = self.http.post(f"{self.url}/v1/models/{self.model}:predict_raw", data=encode_request(dense, ids))
        r.raise_for_status()
        return decode_response(r.content)[0]
