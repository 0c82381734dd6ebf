"""Protobuf PredictRequest / PredictResponse (wire-compatible with the reference's serving/processor/serving/predict.proto) over the
native codec (csrc/common/predict_pb.h, C ABI in csrc/host/predict_codec.cc).  No protoc / generated code: the codec is hand-written
against the wire format, and ``descriptor_pool()`` builds the same messages for google.protobuf at runtime (tests, python clients)."""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from .. import _native

_BOUND = False


def _lib():
    global _BOUND
    L = _native.host()
    if not _BOUND:
        vp, i64 = C.c_void_p, C.c_int64
        L.dr_pb_last_error.restype = C.c_char_p
        L.dr_pb_free.argtypes = [vp]
        L.dr_pb_request_to_wire.restype, L.dr_pb_request_to_wire.argtypes = C.c_int, [C.c_char_p, i64, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(i64)]
        L.dr_pb_response_from_wire.restype = C.c_int
        L.dr_pb_response_from_wire.argtypes = [C.c_char_p, i64, C.c_char_p, i64, C.POINTER(vp), C.POINTER(i64)]
        L.dr_pb_encode_request.restype = C.c_int
        L.dr_pb_encode_request.argtypes = [vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(i64)]
        L.dr_pb_decode_response.restype, L.dr_pb_decode_response.argtypes = i64, [C.c_char_p, i64, vp, i64, C.POINTER(i64)]
        L.dr_pb_response_cols.restype, L.dr_pb_response_cols.argtypes = i64, [C.c_char_p, i64]
        _BOUND = True
    return L


def _take(L, out: C.c_void_p, n: C.c_int64) -> bytes:
    data = C.string_at(out, n.value) if n.value else b""
    L.dr_pb_free(out)
    return data


def encode_predict_request(dense: np.ndarray, ids: np.ndarray, per_feature: bool = False, signature_name: str = "",
                           output_filter: str = "") -> bytes:
    """dense [B, nd] float32, ids [ns, B] int64 -> PredictRequest bytes ({"dense","ids"} or I1../C1.. per-feature inputs)."""
    L = _lib()
    dense = np.ascontiguousarray(dense, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    B, nd = dense.shape
    ns = ids.shape[0]
    assert ids.shape[1] == B
    out, n = C.c_void_p(), C.c_int64()
    rc = L.dr_pb_encode_request(dense.ctypes.data, ids.ctypes.data, B, nd, ns, int(per_feature), signature_name.encode(), output_filter.encode(),
                                C.byref(out), C.byref(n))
    if rc != 0:
        raise RuntimeError("dr_pb_encode_request failed")
    return _take(L, out, n)


def decode_predict_response(pb: bytes) -> Tuple[np.ndarray, int]:
    """PredictResponse bytes -> (probabilities float32 [B] -- [B, num_outputs] for multi-task models --, model_version or -1)."""
    L = _lib()
    ver = C.c_int64(-1)
    cnt = L.dr_pb_decode_response(pb, len(pb), None, 0, C.byref(ver))
    if cnt < 0:
        raise ValueError(L.dr_pb_last_error().decode())
    probs = np.empty(cnt, dtype=np.float32)
    L.dr_pb_decode_response(pb, len(pb), probs.ctypes.data, cnt, C.byref(ver))
    cols = int(L.dr_pb_response_cols(pb, len(pb)))                  # multi-task models answer [B, num_outputs]
    return (probs.reshape(-1, cols) if cols > 1 else probs), int(ver.value)


def request_to_wire(pb: bytes, num_dense: int, num_sparse: int) -> bytes:
    """PredictRequest -> the runtime's compact request (``processor.encode_request`` layout)."""
    L = _lib()
    out, n = C.c_void_p(), C.c_int64()
    if L.dr_pb_request_to_wire(pb, len(pb), num_dense, num_sparse, C.byref(out), C.byref(n)) != 0:
        raise ValueError(L.dr_pb_last_error().decode())
    return _take(L, out, n)


def response_from_wire(wire: bytes, request_pb: bytes = b"") -> bytes:
    """compact response (+ the request, for its output_filter) -> PredictResponse bytes."""
    L = _lib()
    out, n = C.c_void_p(), C.c_int64()
    if L.dr_pb_response_from_wire(wire, len(wire), request_pb or None, len(request_pb), C.byref(out), C.byref(n)) != 0:
        raise ValueError(L.dr_pb_last_error().decode())
    return _take(L, out, n)


def is_wire_request(payload: bytes) -> bool:
    return len(payload) >= 4 and payload[:4] == b"DRRQ"


def message_classes():
    """google.protobuf message classes for the same schema, built at runtime (``tensorflow.eas`` package of predict.proto)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="deeprec_b200/predict.proto", package="tensorflow.eas", syntax="proto3")
    shape = fd.message_type.add(name="ArrayShape")
    shape.field.add(name="dim", number=1, type=F.TYPE_INT64, label=F.LABEL_REPEATED)
    arr = fd.message_type.add(name="ArrayProto")
    arr.field.add(name="dtype", number=1, type=F.TYPE_INT32, label=F.LABEL_OPTIONAL)      # enum on the wire == int32 varint
    arr.field.add(name="array_shape", number=2, type=F.TYPE_MESSAGE, type_name=".tensorflow.eas.ArrayShape", label=F.LABEL_OPTIONAL)
    for name, num, ty in (("float_val", 3, F.TYPE_FLOAT), ("double_val", 4, F.TYPE_DOUBLE), ("int_val", 5, F.TYPE_INT32),
                          ("string_val", 6, F.TYPE_BYTES), ("int64_val", 7, F.TYPE_INT64), ("bool_val", 8, F.TYPE_BOOL)):
        arr.field.add(name=name, number=num, type=ty, label=F.LABEL_REPEATED)

    def add_map(msg, field, number):
        entry = msg.nested_type.add(name="".join(p.capitalize() for p in field.split("_")) + "Entry")
        entry.options.map_entry = True
        entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
        entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, type_name=".tensorflow.eas.ArrayProto", label=F.LABEL_OPTIONAL)
        msg.field.add(name=field, number=number, type=F.TYPE_MESSAGE, label=F.LABEL_REPEATED,
                      type_name=f".tensorflow.eas.{msg.name}.{entry.name}")

    req = fd.message_type.add(name="PredictRequest")
    req.field.add(name="signature_name", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    add_map(req, "inputs", 2)
    req.field.add(name="output_filter", number=3, type=F.TYPE_STRING, label=F.LABEL_REPEATED)
    resp = fd.message_type.add(name="PredictResponse")
    add_map(resp, "outputs", 1)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.eas." + n))
    return get("PredictRequest"), get("PredictResponse"), get("ArrayProto")
