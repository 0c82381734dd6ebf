"""Protobuf PredictRequest / PredictResponse codec (native, hand-written against the wire format) vs google.protobuf as the oracle."""
import numpy as np
import pytest
import torch

from deeprec_b200.serving import predict_pb
from deeprec_b200.serving.processor import decode_response, encode_request


def _batch(B=7, nd=13, ns=26, seed=0):
    rng = np.random.default_rng(seed)
    dense = rng.standard_normal((B, nd)).astype(np.float32)
    ids = rng.integers(-5, 1 << 40, size=(ns, B), dtype=np.int64)      # negative ids exercise 10-byte varints
    return dense, ids


def test_native_encoder_is_parsed_by_google_protobuf():
    Req, Resp, Arr = predict_pb.message_classes()
    dense, ids = _batch()
    for per_feature in (False, True):
        pb = predict_pb.encode_predict_request(dense, ids, per_feature=per_feature, signature_name="serving_default", output_filter="probabilities")
        m = Req.FromString(pb)
        assert m.signature_name == "serving_default" and list(m.output_filter) == ["probabilities"]
        if per_feature:
            assert sorted(m.inputs) == sorted([f"I{i}" for i in range(1, 14)] + [f"C{i}" for i in range(1, 27)])
            assert np.array_equal(np.asarray(m.inputs["C3"].int64_val), ids[2]) and m.inputs["C3"].dtype == 9
            assert np.allclose(np.asarray(m.inputs["I13"].float_val, dtype=np.float32), dense[:, 12]) and list(m.inputs["I13"].array_shape.dim) == [7]
        else:
            assert list(m.inputs["dense"].array_shape.dim) == [7, 13] and list(m.inputs["ids"].array_shape.dim) == [26, 7]
            assert np.array_equal(np.asarray(m.inputs["ids"].int64_val).reshape(26, 7), ids)
            assert np.array_equal(np.asarray(m.inputs["dense"].float_val, dtype=np.float32).reshape(7, 13), dense)


def test_google_protobuf_requests_are_parsed_by_the_native_decoder():
    Req, Resp, Arr = predict_pb.message_classes()
    dense, ids = _batch(B=5)
    want = encode_request(dense, ids)
    # (a) packed tensors, feature-major ids
    m = Req(signature_name="x")
    m.inputs["dense"].dtype = 1; m.inputs["dense"].array_shape.dim.extend([5, 13]); m.inputs["dense"].float_val.extend(dense.reshape(-1).tolist())
    m.inputs["ids"].dtype = 9; m.inputs["ids"].array_shape.dim.extend([26, 5]); m.inputs["ids"].int64_val.extend(ids.reshape(-1).tolist())
    assert predict_pb.request_to_wire(m.SerializeToString(), 13, 26) == want
    # (b) sample-major ids are transposed according to the declared shape
    m2 = Req(); m2.inputs["dense"].CopyFrom(m.inputs["dense"])
    m2.inputs["ids"].dtype = 9; m2.inputs["ids"].array_shape.dim.extend([5, 26]); m2.inputs["ids"].int64_val.extend(ids.T.reshape(-1).tolist())
    assert predict_pb.request_to_wire(m2.SerializeToString(), 13, 26) == want
    # (c) one input per feature, mixed numeric types (double dense column, int32 ids), inserted in shuffled order
    m3 = Req()
    order = np.random.default_rng(1).permutation(39)
    for j in order:
        if j < 13:
            a = m3.inputs[f"I{j + 1}"]
            if j % 2:
                a.dtype = 2; a.double_val.extend(dense[:, j].astype(np.float64).tolist())
            else:
                a.dtype = 1; a.float_val.extend(dense[:, j].tolist())
            a.array_shape.dim.extend([5, 1])
        else:
            a = m3.inputs[f"C{j - 12}"]; a.dtype = 9; a.int64_val.extend(ids[j - 13].tolist())
    assert predict_pb.request_to_wire(m3.SerializeToString(), 13, 26) == want
    small = Req(); small.inputs["I1"].dtype = 1; small.inputs["I1"].float_val.extend([1.0, 2.0])
    small.inputs["C1"].dtype = 3; small.inputs["C1"].int_val.extend([7, -3])
    w = predict_pb.request_to_wire(small.SerializeToString(), 1, 1)
    assert w == encode_request(np.array([[1.0], [2.0]], np.float32), np.array([[7, -3]], np.int64))


def test_malformed_and_mismatched_requests_are_rejected():
    Req, _, _ = predict_pb.message_classes()
    with pytest.raises(ValueError):
        predict_pb.request_to_wire(b"\x12\xff\xff\xff\xff\x0f", 13, 26)                  # truncated length-delimited field
    m = Req(); m.inputs["I1"].dtype = 1; m.inputs["I1"].float_val.extend([1.0])
    with pytest.raises(ValueError, match="expected 13 float and 26 integer inputs"):
        predict_pb.request_to_wire(m.SerializeToString(), 13, 26)
    m = Req(); m.inputs["I1"].dtype = 1; m.inputs["I1"].float_val.extend([1.0, 2.0]); m.inputs["C1"].dtype = 9; m.inputs["C1"].int64_val.extend([1])
    with pytest.raises(ValueError, match="different batch"):
        predict_pb.request_to_wire(m.SerializeToString(), 1, 1)


def test_response_roundtrip_and_output_filter():
    import struct
    Req, Resp, _ = predict_pb.message_classes()
    probs = np.linspace(0.1, 0.9, 6).astype(np.float32)
    wire = struct.pack("<4Iq", 0x53525244, 6, 200, 0, 42) + probs.tobytes()
    assert decode_response(wire)[2] == 42
    pb = predict_pb.response_from_wire(wire)
    r = Resp.FromString(pb)
    assert np.allclose(np.asarray(r.outputs["probabilities"].float_val, dtype=np.float32), probs) and list(r.outputs["model_version"].int64_val) == [42]
    got, ver = predict_pb.decode_predict_response(pb)
    assert np.array_equal(got, probs) and ver == 42
    # a response produced by google.protobuf is decoded by the native client path too
    g = Resp(); g.outputs["probabilities"].dtype = 1; g.outputs["probabilities"].float_val.extend(probs.tolist())
    got, ver = predict_pb.decode_predict_response(g.SerializeToString())
    assert np.allclose(got, probs) and ver == -1
    flt = Req(output_filter=["probabilities"]).SerializeToString()
    assert list(Resp.FromString(predict_pb.response_from_wire(wire, flt)).outputs) == ["probabilities"]


def test_http_predict_proto_endpoint():
    from starlette.testclient import TestClient
    import deeprec_b200 as dr
    from deeprec_b200.models.zoo import build_model
    from deeprec_b200.serving import SessionGroup
    from deeprec_b200.serving.http_server import HttpClient, ServingBackend, create_app
    torch.manual_seed(0)
    model = build_model("deepfm", device="cpu")
    g = torch.Generator().manual_seed(1)
    dense = torch.randn(16, 13, generator=g); ids = torch.randint(0, 50, (26, 16), generator=g)
    group = SessionGroup(model, session_num=1)
    ref = torch.sigmoid(group.run(dense, ids)).numpy()
    app = create_app({"deepfm": ServingBackend.from_session_group(group, version=5, extra_info={"num_dense": 13, "num_sparse": 26})})
    with TestClient(app) as http:
        cli = HttpClient("http://testserver", "deepfm", session=http)
        assert np.allclose(cli.predict_proto(dense.numpy(), ids.numpy()), ref, atol=1e-6)
        assert np.allclose(cli.predict_proto(dense.numpy(), ids.numpy(), per_feature=True), ref, atol=1e-6)
        assert http.post("/v1/models/deepfm:predict_proto", content=b"\x12\xff\xff\xff\xff\x0f").status_code == 400
