"""gRPC front-end for the Processors (the reference's processor sits behind a gRPC PredictionService in EAS / its TF-Serving fork).

  service tensorflow.eas.PredictService {
    rpc Predict(PredictRequest) returns (PredictResponse);          // messages of serving/processor/serving/predict.proto
    rpc GetModelInfo(ServingModelInfoRequest) returns (ServingModelInfo-as-JSON bytes);
  }

No generated stubs: the request / response bytes ARE what ``process()`` consumes and produces (the native runtimes parse the protobuf
themselves), so the handlers are registered as generic byte-level methods.  Any client generated from predict.proto -- or the
``PredictClient`` below -- can call it.  One server can host several models: the model is chosen by the ``model`` metadata key (default:
the only / first model).
"""
from __future__ import annotations

import json
from concurrent import futures
from typing import Dict, Optional

import numpy as np

SERVICE = "tensorflow.eas.PredictService"


def _identity(b):
    return b


def create_server(processors: Dict[str, object], address: str = "127.0.0.1:0", max_workers: int = 8):
    """``processors``: {model name: Processor | ProcessorGroup | anything with ``process(bytes) -> (rc, bytes)`` and ``model_info()``}.
    Returns ``(grpc.Server, bound_port)`` (already started)."""
    import grpc
    default = next(iter(processors))

    def pick(context):
        name = dict(context.invocation_metadata()).get("model", default)
        proc = processors.get(name)
        if proc is None:
            context.abort(grpc.StatusCode.NOT_FOUND, f"unknown model {name!r}; serving: {sorted(processors)}")
        return proc

    def predict(request: bytes, context):
        rc, out = pick(context).process(request)
        if rc != 200:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT if rc == 500 else grpc.StatusCode.INTERNAL, f"process returned {rc}")
        return out

    def model_info(request: bytes, context):
        return json.dumps(pick(context).model_info()).encode()

    handler = grpc.method_handlers_generic_handler(SERVICE, {
        "Predict": grpc.unary_unary_rpc_method_handler(predict, request_deserializer=_identity, response_serializer=_identity),
        "GetModelInfo": grpc.unary_unary_rpc_method_handler(model_info, request_deserializer=_identity, response_serializer=_identity),
    })
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((handler,))
    port = server.add_insecure_port(address)
    server.start()
    return server, port


class PredictClient:
    """Minimal gRPC client (serving/sdk python demo): numpy in, probabilities out; requests are real ``PredictRequest`` protobufs."""

    def __init__(self, target: str, model: Optional[str] = None):
        import grpc
        self.channel = grpc.insecure_channel(target)
        self._predict = self.channel.unary_unary(f"/{SERVICE}/Predict", request_serializer=_identity, response_deserializer=_identity)
        self._info = self.channel.unary_unary(f"/{SERVICE}/GetModelInfo", request_serializer=_identity, response_deserializer=_identity)
        self.metadata = (("model", model),) if model else None

    def predict(self, dense: np.ndarray, ids: np.ndarray, per_feature: bool = False, timeout: float = 10.0):
        from .predict_pb import decode_predict_response, encode_predict_request
        out = self._predict(encode_predict_request(dense, ids, per_feature=per_feature), metadata=self.metadata, timeout=timeout)
        return decode_predict_response(out)          # (probabilities, model_version)

    def predict_raw(self, request_pb: bytes, timeout: float = 10.0) -> bytes:
        return self._predict(request_pb, metadata=self.metadata, timeout=timeout)

    def model_info(self, timeout: float = 10.0) -> dict:
        return json.loads(self._info(b"", metadata=self.metadata, timeout=timeout).decode())

    def close(self) -> None:
        self.channel.close()
