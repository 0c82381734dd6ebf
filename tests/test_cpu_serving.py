"""Native CPU Processor (csrc/host/cpu_serving.cc) end to end: train the DLRM module on CPU -> export -> initialize -> process /
batch_process (compact and protobuf encodings) == the module's own predictions; delta update patches the live model; full update swaps
after warm-up; invalid versions are skipped; HTTP front-end on top."""
import os
import time

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving import (Processor, decode_response, encode_request, export_delta_module, export_saved_model_module, predict_pb)

CARDS = [50, 1000, 7, 300] + [97] * 22


def _train(model, opt, steps, seed):
    for s in range(steps):
        d, ids, y = criteo_batch(512, 13, CARDS, seed=seed + s)
        loss = model.loss(d, ids, y); opt.zero_grad(); loss.backward(); opt.step()
    return d, ids


def _ref(model, d, ids):
    model.eval()
    with torch.no_grad():
        p = torch.sigmoid(model(d, ids)).numpy().copy()
    model.train()
    return p


def _wait(pred, timeout=20.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.05)
    return False


def test_cpu_processor_matches_module_and_hot_swaps(tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 6, 0)
    root = str(tmp_path)
    export_saved_model_module(model, os.path.join(root, "v1"), version=6, root=root)
    proc = Processor(os.path.join(root, "v1"), {"session_num": 3, "select_session_policy": "RR", "max_batch": 200, "checkpoint_dir": root,
                                                "model_update_interval_ms": 100, "timeline_interval_step": 1, "timeline_start_step": 0,
                                                "timeline_trace_count": 3, "timeline_path": os.path.join(root, "trace.jsonl")}, device="cpu")
    ref = _ref(model, d, ids)
    got = proc.predict(d.numpy(), ids.numpy())                         # 512 rows > max_batch 200 -> chunked inside the session
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-5, np.abs(got - ref).max()
    # unseen ids read default rows exactly like the module
    ids2 = ids.clone(); ids2[:, :50] += 10 ** 9
    assert np.abs(proc.predict(d.numpy(), ids2.numpy()) - _ref(model, d, ids2)).max() < 1e-5
    rc, outs = proc.batch_process([encode_request(d.numpy()[:64], ids.numpy()[:, :64]) for _ in range(4)])
    assert rc == 200 and all(np.abs(decode_response(o)[0] - ref[:64]).max() < 1e-5 for o in outs)
    # protobuf PredictRequest (both input conventions) straight into process()
    for per_feature in (False, True):
        probs, version = predict_pb.decode_predict_response(proc.predict_proto(predict_pb.encode_predict_request(d.numpy(), ids.numpy(), per_feature=per_feature)))
        assert version == 6 and np.array_equal(probs, got)
    assert proc.process(b"garbage")[0] == 500 and proc.process(encode_request(d.numpy()[:, :5], ids.numpy()))[0] == 500
    info = proc.model_info()
    assert info["model_version"] == 6 and info["sessions"] == 3 and info["device"] == "cpu" and info["requests"] >= 8 and info["failures"] >= 1
    assert len(open(os.path.join(root, "trace.jsonl")).read().strip().splitlines()) == 3
    # ---- delta update: only the touched rows + the dense block travel; the live model picks them up without a swap
    _train(model, opt, 3, 100)
    export_delta_module(model, root, base_version=6, version=9)
    ref2 = _ref(model, d, ids)
    assert _wait(lambda: proc.model_info()["delta_updates"] >= 1)
    got2 = proc.predict(d.numpy(), ids.numpy())
    assert proc.model_info()["delta_version"] == 9 and np.abs(got2 - ref2).max() < 1e-5 and np.abs(ref2 - ref).max() > 1e-4
    # ---- an invalid full version is skipped, the old model keeps serving
    os.makedirs(os.path.join(root, "bad"))
    open(os.path.join(root, "bad", "saved_model.json"), "w").write("{not json")
    from deeprec_b200.serving.export import _write_versions
    _write_versions(root, full={"version": 10, "dir": os.path.join(root, "bad")})
    time.sleep(0.5)
    assert proc.model_info()["model_version"] == 6 and np.abs(proc.predict(d.numpy(), ids.numpy()) - ref2).max() < 1e-5
    # ---- full update: new version directory, swapped after warm-up
    _train(model, opt, 2, 200)
    export_saved_model_module(model, os.path.join(root, "v2"), version=11, root=root)
    assert _wait(lambda: proc.model_info()["model_version"] == 11)
    ref3 = _ref(model, d, ids)
    assert np.abs(proc.predict(d.numpy(), ids.numpy()) - ref3).max() < 1e-5 and proc.model_info()["full_updates"] == 1
    proc.close()


def test_cpu_processor_concurrent_requests_and_http(tmp_path):
    import threading
    from starlette.testclient import TestClient
    from deeprec_b200.serving.http_server import HttpClient, ServingBackend, create_app
    dr.embedding_variable.clear_registry()
    torch.manual_seed(1)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 3, 7)
    export_saved_model_module(model, str(tmp_path / "m"), version=3)
    proc = Processor(str(tmp_path / "m"), {"session_num": 4, "select_session_policy": "MOD", "max_batch": 512, "model_update_interval_ms": 0}, device="cpu")
    ref = _ref(model, d, ids)
    errs = []

    def client(i):
        try:
            for k in range(10):
                lo = (i * 17 + k * 31) % 400
                out = proc.predict(d.numpy()[lo:lo + 64], ids.numpy()[:, lo:lo + 64])
                assert np.abs(out - ref[lo:lo + 64]).max() < 1e-5
        except Exception as e:      # pragma: no cover
            errs.append(e)
    ts = [threading.Thread(target=client, args=(i,)) for i in range(8)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs and proc.model_info()["requests"] == 80
    with TestClient(create_app({"dlrm": ServingBackend.from_processor(proc)})) as http:
        cli = HttpClient("http://testserver", "dlrm", session=http)
        assert np.abs(cli.predict(d.numpy()[:8], ids.numpy()[:, :8]) - ref[:8]).max() < 1e-5
        assert np.abs(cli.predict_proto(d.numpy()[:8], ids.numpy()[:, :8], per_feature=True) - ref[:8]).max() < 1e-5
        assert http.get("/v1/models/dlrm").json()["device"] == "cpu"
    proc.close()
    with pytest.raises(RuntimeError):
        Processor(str(tmp_path / "nope"), {}, device="cpu")


def test_cpu_processor_with_redis_feature_store(tmp_path):
    """feature_store_type = redis: embedding rows live in the (mini) Redis server, the processor keeps the dense net and default rows;
    answers equal the local-mode processor and the module, a new full version reads its own key space, an unreachable store is a 500."""
    from deeprec_b200.serving.feature_store import MiniRedisServer, RedisFeatureStore, export_processor_tables
    dr.embedding_variable.clear_registry()
    torch.manual_seed(2)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 4, 3)
    root = str(tmp_path)
    srv = MiniRedisServer()
    store = RedisFeatureStore(srv.host, srv.port)
    evs = model.embedding_variables()
    n = export_processor_tables(evs, store, "ctr_model", 4)
    assert n == sum(e.total_count() for e in evs) == store.dbsize()
    export_saved_model_module(model, os.path.join(root, "v1"), version=4, root=root)
    cfg = {"session_num": 2, "max_batch": 256, "checkpoint_dir": root, "model_update_interval_ms": 100}
    remote = Processor(os.path.join(root, "v1"), dict(cfg, feature_store_type="redis", redis_url=f"{srv.host}:{srv.port}", redis_prefix="ctr_model"), device="cpu")
    local = Processor(os.path.join(root, "v1"), dict(cfg, model_update_interval_ms=0), device="cpu")
    ids2 = ids.clone(); ids2[:, :40] += 10 ** 9                       # ids the store has never seen -> default rows
    ref = _ref(model, d, ids2)
    a, b = remote.predict(d.numpy(), ids2.numpy()), local.predict(d.numpy(), ids2.numpy())
    assert np.abs(a - ref).max() < 1e-5 and np.array_equal(a, b) and remote.model_info()["feature_store_type"] == "redis"
    # new full version: rows first (own key space), then the saved model; the processor swaps and reads version 7 rows
    _train(model, opt, 3, 50)
    export_processor_tables(evs, store, "ctr_model", 7)
    export_saved_model_module(model, os.path.join(root, "v2"), version=7, root=root)
    assert _wait(lambda: remote.model_info()["model_version"] == 7)
    assert np.abs(remote.predict(d.numpy(), ids2.numpy()) - _ref(model, d, ids2)).max() < 1e-5
    # store goes away -> requests fail cleanly (500), the process survives
    srv.close(); store.close()
    rc, _ = remote.process(encode_request(d.numpy()[:8], ids.numpy()[:, :8]))
    assert rc == 500 and remote.model_info()["failures"] >= 1
    # ... and comes back (same port): the session reconnects on the next request
    srv2 = MiniRedisServer(port=srv.port)
    store2 = RedisFeatureStore(srv2.host, srv2.port)
    export_processor_tables(evs, store2, "ctr_model", 7)
    assert np.abs(remote.predict(d.numpy(), ids2.numpy()) - _ref(model, d, ids2)).max() < 1e-5
    store2.close(); srv2.close()
    remote.close(); local.close()
    with pytest.raises(RuntimeError):
        Processor(os.path.join(root, "v2"), dict(cfg, feature_store_type="redis", redis_url="127.0.0.1:1"), device="cpu")


def test_c_client_integrates_the_abi(tmp_path):
    """examples/c_client.c: a plain C program dlopens the runtime, initializes an exported model and sends protobuf + compact requests."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dr.embedding_variable.clear_registry()
    torch.manual_seed(3)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    _train(model, opt, 2, 11)
    export_saved_model_module(model, str(tmp_path / "m"), version=2)
    exe = str(tmp_path / "c_client")
    b = subprocess.run(["gcc", "-std=c99", "-Wall", os.path.join(root, "examples", "c_client.c"), "-I" + os.path.join(root, "deeprec_b200", "csrc", "include"), "-ldl", "-o", exe],
                       capture_output=True, text=True)
    assert b.returncode == 0, b.stderr
    from deeprec_b200 import build as _b
    r = subprocess.run([exe, os.path.join(_b.LIB, "libdeeprec_host.so"), str(tmp_path / "m")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "C_CLIENT_OK" in r.stdout and "model version 2" in r.stdout, r.stdout + r.stderr


def test_grpc_front_end_over_the_cpu_processor(tmp_path):
    """gRPC PredictService (byte-level handlers, no generated stubs): real PredictRequest protobufs in, PredictResponse out; two models on
    one server selected by metadata; malformed requests map to INVALID_ARGUMENT; a google.protobuf-built request works too."""
    import grpc
    from deeprec_b200.serving.grpc_server import PredictClient, create_server
    procs = {}
    refs = {}
    for name, seed in (("ctr", 5), ("cvr", 6)):
        dr.embedding_variable.clear_registry()
        torch.manual_seed(seed)
        model = build_model("dlrm", device="cpu", cardinalities=CARDS)
        opt = dr.optim.AdagradOptimizer(model, lr=0.05)
        _train(model, opt, 2, seed)
        d, ids, _ = criteo_batch(512, 13, CARDS, seed=99)
        export_saved_model_module(model, str(tmp_path / name), version=seed)
        procs[name] = Processor(str(tmp_path / name), {"session_num": 2, "model_update_interval_ms": 0}, device="cpu")
        refs[name] = _ref(model, d, ids)
    server, port = create_server(procs)
    try:
        for name in ("ctr", "cvr"):
            cli = PredictClient(f"127.0.0.1:{port}", model=name)
            probs, version = cli.predict(d.numpy(), ids.numpy())
            assert version == {"ctr": 5, "cvr": 6}[name] and np.abs(probs - refs[name]).max() < 1e-5
            assert np.array_equal(cli.predict(d.numpy(), ids.numpy(), per_feature=True)[0], probs)
            assert cli.model_info()["device"] == "cpu"
            cli.close()
        cli = PredictClient(f"127.0.0.1:{port}")                      # no metadata -> the first model
        Req, Resp, _ = predict_pb.message_classes()
        m = Req(output_filter=["probabilities"])
        m.inputs["dense"].dtype = 1; m.inputs["dense"].float_val.extend(d.numpy()[:4].reshape(-1).tolist())
        m.inputs["ids"].dtype = 9; m.inputs["ids"].int64_val.extend(ids.numpy()[:, :4].reshape(-1).tolist())
        r = Resp.FromString(cli.predict_raw(m.SerializeToString()))
        assert list(r.outputs) == ["probabilities"] and np.abs(np.asarray(r.outputs["probabilities"].float_val) - refs["ctr"][:4]).max() < 1e-5
        with pytest.raises(grpc.RpcError) as e:
            cli.predict_raw(b"\x12\xff\xff\xff\xff\x0f")
        assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
        bad = PredictClient(f"127.0.0.1:{port}", model="nope")
        with pytest.raises(grpc.RpcError) as e:
            bad.model_info()
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
        bad.close(); cli.close()
    finally:
        server.stop(0)
        for p in procs.values():
            p.close()


def test_model_server_cli_serves_grpc_and_http(tmp_path):
    """``python -m deeprec_b200.serving.serve``: a separate process loads the export, listens on gRPC + HTTP, answers both, exits on SIGTERM."""
    import json
    import signal
    import subprocess
    import sys

    import requests

    from deeprec_b200.serving.grpc_server import PredictClient
    from deeprec_b200.serving.http_server import HttpClient
    dr.embedding_variable.clear_registry()
    torch.manual_seed(3)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 2, 3)
    export_saved_model_module(model, str(tmp_path / "m"), version=9)
    ref = _ref(model, d, ids)
    pf = str(tmp_path / "ports.json")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    p = subprocess.Popen([sys.executable, "-m", "deeprec_b200.serving.serve", "--model", f"ctr={tmp_path / 'm'}", "--device", "cpu", "--config",
                          '{"session_num": 2, "model_update_interval_ms": 0}', "--grpc_port", "0", "--http_port", "0", "--port_file", pf], env=env)
    try:
        assert _wait(lambda: os.path.exists(pf) or p.poll() is not None, 120) and p.poll() is None
        ports = json.load(open(pf))
        cli = PredictClient(f"127.0.0.1:{ports['grpc']}", model="ctr")
        probs, version = cli.predict(d.numpy(), ids.numpy(), timeout=60)
        assert version == 9 and np.abs(probs - ref).max() < 1e-5
        cli.close()
        base = f"http://127.0.0.1:{ports['http']}"
        assert _wait(lambda: _ok(lambda: requests.get(base + "/healthz", timeout=2).status_code == 200), 30)
        assert np.abs(HttpClient(base, "ctr").predict_proto(d.numpy(), ids.numpy()) - ref).max() < 1e-5
    finally:
        p.send_signal(signal.SIGTERM)
        try:
            assert p.wait(30) == 0
        except subprocess.TimeoutExpired:
            p.kill(); raise


def _ok(fn):
    try:
        return fn()
    except Exception:
        return False


def test_session_cpusets_place_sessions_and_leave_the_caller_alone(tmp_path):
    """``cpusets`` (SessionGroup.md): session i runs on its own CPUs -- observed through sched_getcpu() sampled inside the request -- the
    per-session thread budget follows the set size, and the calling thread gets its own affinity mask back after every request."""
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 3:
        pytest.skip("needs 3 CPUs")
    dr.embedding_variable.clear_registry()
    torch.manual_seed(1)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 2, 1)
    export_saved_model_module(model, str(tmp_path / "m"), version=1)
    ref = _ref(model, d, ids)
    a, b, c = allowed[0], allowed[1], allowed[2]
    proc = Processor(str(tmp_path / "m"), {"session_num": 2, "select_session_policy": "RR", "model_update_interval_ms": 0,
                                           "cpusets": f"{a};{b}-{c}" if c == b + 1 else f"{a};{b},{c}"}, device="cpu")
    try:
        before = os.sched_getaffinity(0)
        for _ in range(4):                                              # RR: both sessions serve, small and team-sized batches
            assert np.abs(proc.predict(d.numpy(), ids.numpy()) - ref).max() < 1e-5
            assert np.abs(proc.predict(d.numpy()[:8], ids.numpy()[:, :8]) - ref[:8]).max() < 1e-5
        assert os.sched_getaffinity(0) == before
        info = proc.model_info()
        assert info["cpusets"] == f"{a};{b},{c}"
        assert info["session_last_cpu"][0] == a and info["session_last_cpu"][1] in (b, c)
    finally:
        proc.close()


def test_request_batching_merges_concurrent_requests(tmp_path):
    """enable_batching: concurrent small requests are merged into one forward pass (leader / follower); every caller gets exactly its own
    rows back, large requests bypass the batcher, and the merge counters show that merging happened."""
    import threading
    dr.embedding_variable.clear_registry()
    torch.manual_seed(2)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 2, 2)
    export_saved_model_module(model, str(tmp_path / "m"), version=4)
    ref = _ref(model, d, ids)
    # one session: whoever arrives while it is busy queues up behind a leader and shares the next forward pass (adaptive mode runs a
    # lone request immediately when a session is idle)
    proc = Processor(str(tmp_path / "m"), {"session_num": 1, "model_update_interval_ms": 0, "enable_batching": True,
                                           "batching_parameters": {"max_batch_size": 16, "batch_timeout_micros": 2000}}, device="cpu")
    try:
        dn, idn = d.numpy(), ids.numpy()
        errs, N, T = [], 24, 8

        def client(t):
            try:
                for i in range(N):
                    r0 = (t * N + i) * 2 % 500
                    rows = 1 + (i % 3)                                   # 1..3 rows per request
                    got = proc.predict(dn[r0:r0 + rows], idn[:, r0:r0 + rows])
                    if got.shape != (rows,) or np.abs(got - ref[r0:r0 + rows]).max() > 1e-5:
                        errs.append((t, i, got, ref[r0:r0 + rows]))
            except Exception as e:  # noqa: BLE001
                errs.append(repr(e))

        th = [threading.Thread(target=client, args=(t,)) for t in range(T)]
        [x.start() for x in th]; [x.join() for x in th]
        assert not errs, errs[:2]
        assert np.abs(proc.predict(dn, idn) - ref).max() < 1e-5           # 512 rows: not batched
        # protobuf requests go through the same batcher
        rc, out = proc.process(predict_pb.encode_predict_request(dn[:2], idn[:, :2]))
        assert rc == 200 and np.abs(predict_pb.decode_predict_response(out)[0] - ref[:2]).max() < 1e-5
        b = proc.model_info()["batching"]
        assert b["max_batch_size"] == 16 and b["merged_requests"] == T * N + 1
        assert b["merged_batches"] < b["merged_requests"], b             # at least some requests shared a forward pass
    finally:
        proc.close()


@pytest.mark.parametrize("name", ["wdl", "deepfm", "dcn", "dcnv2", "masknet"])
def test_op_program_models_on_the_cpu_processor(tmp_path, name):
    """DeepFM / DCN exported as an op program (BatchNorm folded at export): the native CPU Processor reproduces the module's predictions,
    takes a delta update (rows + re-folded dense tensors) and serves protobuf requests; a program with a wrong shape is rejected."""
    import json
    from deeprec_b200.serving import export_delta_program, export_saved_model_program
    dr.embedding_variable.clear_registry()
    torch.manual_seed(4)
    model = build_model(name, device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 4, 4)                                     # BatchNorm statistics move away from (0, 1)
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=4, root=root)
    meta = json.load(open(os.path.join(root, "v1", "saved_model.json")))
    assert meta["arch"] == "program" and meta["model"] == name and {o["op"] for o in meta["program"]} >= {"concat", "linear"}
    proc = Processor(os.path.join(root, "v1"), {"session_num": 2, "max_batch": 200, "checkpoint_dir": root, "model_update_interval_ms": 100}, device="cpu")
    try:
        ref = _ref(model, d, ids)
        got = proc.predict(d.numpy(), ids.numpy())                        # 512 rows > max_batch: chunked, team-sized chunks
        assert np.abs(got - ref).max() < 2e-5, np.abs(got - ref).max()
        assert np.abs(proc.predict(d.numpy()[:3], ids.numpy()[:, :3]) - ref[:3]).max() < 2e-5
        ids2 = ids.clone(); ids2[:, :40] += 10 ** 9                        # unseen ids read default rows like the module
        assert np.abs(proc.predict(d.numpy(), ids2.numpy()) - _ref(model, d, ids2)).max() < 2e-5
        rc, out = proc.process(predict_pb.encode_predict_request(d.numpy()[:5], ids.numpy()[:, :5], per_feature=True))
        assert rc == 200 and np.abs(predict_pb.decode_predict_response(out)[0] - ref[:5]).max() < 2e-5
        assert proc.model_info()["model"] == name
        _train(model, opt, 2, 50)
        export_delta_program(model, root, base_version=4, version=6)
        assert _wait(lambda: proc.model_info()["delta_version"] == 6)
        ref2 = _ref(model, d, ids)
        assert np.abs(ref2 - ref).max() > 1e-4 and np.abs(proc.predict(d.numpy(), ids.numpy()) - ref2).max() < 2e-5
    finally:
        proc.close()
    # request batching merges request-shaped id blocks (WDL: 26 id rows feed 52 tables)
    import threading
    bp = Processor(os.path.join(root, "v1"), {"session_num": 1, "model_update_interval_ms": 0, "enable_batching": True, "max_batch_size": 16,
                                              "batch_timeout_micros": 5000}, device="cpu")
    try:
        base = bp.predict(d.numpy(), ids.numpy())
        bad = []

        def cl(t):
            for i in range(6):
                r0 = (t * 6 + i) * 3
                if np.abs(bp.predict(d.numpy()[r0:r0 + 2], ids.numpy()[:, r0:r0 + 2]) - base[r0:r0 + 2]).max() > 1e-5:
                    bad.append((t, i))
        th = [threading.Thread(target=cl, args=(t,)) for t in range(4)]
        [x.start() for x in th]; [x.join() for x in th]
        assert not bad, bad
    finally:
        bp.close()
    # a corrupted program (output buffer that does not exist) must fail initialisation, not crash
    meta["output"] = "nope"
    json.dump(meta, open(os.path.join(root, "v1", "saved_model.json"), "w"))
    with pytest.raises(RuntimeError):
        Processor(os.path.join(root, "v1"), {"session_num": 1, "model_update_interval_ms": 0}, device="cpu")


def test_full_update_with_a_wider_architecture_resizes_the_sessions(tmp_path):
    """A hot full update may publish a wider model (more tables, more dense columns, wider layers): the session scratch buffers grow under the
    session mutex before the version is published (they were sized from the first model only -> heap overflow)."""
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    cards_a, cards_b = CARDS[:6], CARDS
    from deeprec_b200.models.dlrm import DLRM
    small = DLRM(13, cards_a, 16, mlp_bot=(64, 16), mlp_top=(32,), device="cpu")
    root = str(tmp_path)
    export_saved_model_module(small, os.path.join(root, "v1"), version=1, root=root)
    proc = Processor(os.path.join(root, "v1"), {"session_num": 2, "max_batch": 256, "checkpoint_dir": root, "model_update_interval_ms": 50}, device="cpu")
    d, ids, _ = criteo_batch(256, 13, cards_a, seed=3)
    assert np.abs(proc.predict(d.numpy(), ids.numpy()) - _ref(small, d, ids)).max() < 1e-5
    dr.embedding_variable.clear_registry()
    big = build_model("dlrm", device="cpu", cardinalities=cards_b)          # 26 tables, 512-256-64-16 / 512-256
    export_saved_model_module(big, os.path.join(root, "v2"), version=2, root=root)
    assert _wait(lambda: proc.model_info()["model_version"] == 2)
    d2, ids2, _ = criteo_batch(256, 13, cards_b, seed=4)
    for _ in range(3):                                                     # every session runs the wide model at full batch
        assert np.abs(proc.predict(d2.numpy(), ids2.numpy()) - _ref(big, d2, ids2)).max() < 1e-5
    proc.close()


def test_din_op_program_on_the_cpu_processor(tmp_path):
    """DIN as an op program over lookup COLUMNS that share tables (``col_table``: target item + 20 history positions read the item table):
    valid_mask / seq_zip / seq_mask / seq_sum / din_attention / prelu reproduce the module, padding ids (-1) are masked, a history without
    any valid position gives the zero attention vector, protobuf requests and a delta update work."""
    import json
    from deeprec_b200.data import taobao_batch
    from deeprec_b200.models.rec_engine import din_ids
    from deeprec_b200.serving import export_delta_program, export_saved_model_program
    dr.embedding_variable.clear_registry()
    torch.manual_seed(3)
    L = 20
    model = build_model("din", device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)

    def step(seed):
        b = taobao_batch(256, L, 500, 3000, 40, seed=seed)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
        return b
    for sd in range(4):
        b = step(sd)
    b["hist_item"][:5] = -1; b["hist_cat"][:5] = -1                     # five samples with an empty history
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=2, root=root, max_len=L)
    meta = json.load(open(os.path.join(root, "v1", "saved_model.json")))
    assert meta["col_table"] == [0, 1, 2] + [1] * L + [2] * L and meta["num_tables"] == 3 and meta["num_id_rows"] == 3 + 2 * L
    assert {o["op"] for o in meta["program"]} >= {"valid_mask", "seq_zip", "seq_mask", "seq_sum", "din_attention", "prelu", "linear"}

    def ref():
        model.eval()
        with torch.no_grad():
            p = torch.sigmoid(model(b)).numpy().copy()
        model.train()
        return p
    ids = din_ids(b).numpy(); dense = np.zeros((256, 1), np.float32)
    proc = Processor(os.path.join(root, "v1"), {"session_num": 2, "max_batch": 100, "checkpoint_dir": root, "model_update_interval_ms": 100}, device="cpu")
    try:
        r0 = ref()
        got = proc.predict(dense, ids)                                   # 256 rows > max_batch: chunked
        assert np.abs(got - r0).max() < 2e-5, np.abs(got - r0).max()
        assert np.abs(proc.predict(dense[:3], ids[:, :3]) - r0[:3]).max() < 2e-5
        rc, out = proc.process(predict_pb.encode_predict_request(dense[:7], ids[:, :7]))
        assert rc == 200 and np.abs(predict_pb.decode_predict_response(out)[0] - r0[:7]).max() < 2e-5
        for sd in range(10, 13):
            step(sd)
        export_delta_program(model, root, base_version=2, version=3, max_len=L)
        assert _wait(lambda: proc.model_info()["delta_version"] == 3)
        r1 = ref()
        assert np.abs(r1 - r0).max() > 1e-4 and np.abs(proc.predict(dense, ids) - r1).max() < 2e-5
    finally:
        proc.close()


@pytest.mark.parametrize("name", ["esmm", "mmoe", "dbmtl", "ple", "simple_multitask", "dssm"])
def test_multitask_and_dssm_op_programs_on_the_cpu_processor(tmp_path, name):
    """Taobao-shaped multi-task models (two probabilities per row: ctr, cvr) and the two-tower DSSM as op programs: softmax-gated expert
    mixtures (seq_mask + seq_sum), cosine similarity, multi-output responses in the compact AND the protobuf encoding."""
    import json
    from deeprec_b200.data import taobao_batch
    from deeprec_b200.models.rec_engine import din_ids
    from deeprec_b200.serving import export_saved_model_program
    dr.embedding_variable.clear_registry()
    torch.manual_seed(5)
    L = 12
    model = build_model(name, device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    for sd in range(3):
        b = taobao_batch(128, L, 500, 3000, 40, seed=sd)
        loss = model.loss(b); opt.zero_grad(); loss.backward(); opt.step()
    b["hist_item"][:4] = -1; b["hist_cat"][:4] = -1
    root = str(tmp_path)
    export_saved_model_program(model, os.path.join(root, "v1"), version=1, root=root, max_len=L)
    meta = json.load(open(os.path.join(root, "v1", "saved_model.json")))
    model.eval()
    with torch.no_grad():
        out = model(b)
    if name == "dssm":
        ref = torch.sigmoid(out).numpy()
        assert "num_outputs" not in meta and "cosine" in {o["op"] for o in meta["program"]}
    else:
        ref = torch.stack([torch.sigmoid(out["ctr"]), torch.sigmoid(out["cvr"])], 1).numpy()
        assert meta["num_outputs"] == 2 and meta["output_names"] == ["ctr", "cvr"]
    ids = din_ids(b).numpy(); dense = np.zeros((128, 1), np.float32)
    proc = Processor(os.path.join(root, "v1"), {"session_num": 2, "max_batch": 50, "model_update_interval_ms": 0}, device="cpu")
    try:
        got = proc.predict(dense, ids)                                   # 128 rows > max_batch: chunked
        assert got.shape == ref.shape and np.abs(got - ref).max() < 3e-5, np.abs(got - ref).max()
        rc, pb = proc.process(predict_pb.encode_predict_request(dense[:5], ids[:, :5]))
        assert rc == 200
        dec = predict_pb.decode_predict_response(pb)[0]
        assert dec.shape == ref[:5].shape and np.abs(dec - ref[:5]).max() < 3e-5
    finally:
        proc.close()


def test_the_reference_model_config_is_accepted_key_by_key(tmp_path):
    """The JSON a DeepRec Processor deployment already has (docs/docs_en/Processor.md "Configure file", serving/model_config.cc) initialises both native
    runtimes unchanged: implemented keys act, mapped keys map (omp_num_threads -> session team, model_update_intra_threads, gpu_ids_list), structural keys
    are reported, what this build cannot honour (oss / hdfs stores, another serialize_protocol) fails initialisation with the reason, typos are reported."""
    dr.embedding_variable.clear_registry()
    torch.manual_seed(2)
    model = build_model("dlrm", device="cpu", cardinalities=CARDS)
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    d, ids = _train(model, opt, 2, 2)
    root = str(tmp_path)
    export_saved_model_module(model, os.path.join(root, "v1"), version=2, root=root)
    ref = _ref(model, d, ids)
    reference_config = {
        "session_num": 2, "select_session_policy": "MOD", "use_per_session_threads": False, "gpu_ids_list": "0,2", "use_multi_stream": False,
        "enable_device_placement_optimization": False, "enable_inline_execute": False, "omp_num_threads": 2, "kmp_blocktime": 0, "feature_store_type": "local",
        "read_thread_num": 4, "update_thread_num": 1, "serialize_protocol": "protobuf", "inter_op_parallelism_threads": 10, "model_update_inter_threads": 4,
        "model_update_intra_threads": 2, "init_timeout_minutes": 1, "signature_name": "serving_default", "model_store_type": "local",
        "checkpoint_dir": root, "savedmodel_dir": os.path.join(root, "v1"), "timeline_start_step": 1, "timeline_interval_step": 2, "timeline_trace_count": 3,
        "model_update_interval_ms": 100, "sesion_num": 7}                                  # <- a typo: must be reported, not silently defaulted
    for device in ("cpu", "cuda_emu"):
        proc = Processor(os.path.join(root, "v1"), dict(reference_config, max_batch=64 if device != "cpu" else 4096), device=device)
        try:
            n = 64 if device != "cpu" else 512
            assert np.abs(proc.predict(d.numpy()[:n], ids.numpy()[:, :n]) - ref[:n]).max() < (1e-5 if device == "cpu" else 3e-2)
            info = proc.model_info()
            mc = info["model_config"]
            assert mc["unknown"] == ["sesion_num"] and mc["signature_name"] == "serving_default"
            assert {"use_per_session_threads", "use_multi_stream", "inter_op_parallelism_threads", "model_update_inter_threads", "kmp_blocktime"} <= set(mc["structural"])
            assert info["sessions"] == 2
            if device == "cpu":
                assert info["threads_per_session"] == 2 and mc["model_update_intra_threads"] == 2           # omp_num_threads -> the session's OpenMP team
            else:
                assert info["gpu_id"] == 0                                                                  # first entry of gpu_ids_list
        finally:
            proc.close()
    # a hot update still lands with model_update_intra_threads bounding the updater's team
    proc = Processor(os.path.join(root, "v1"), reference_config, device="cpu")
    try:
        _train(model, opt, 2, 9)
        export_delta_module(model, root, base_version=2, version=4)
        assert _wait(lambda: proc.model_info()["delta_version"] == 4)
        assert np.abs(proc.predict(d.numpy(), ids.numpy()) - _ref(model, d, ids)).max() < 1e-5
    finally:
        proc.close()
    for bad in ({"model_store_type": "oss"}, {"checkpoint_dir": "oss://bucket/ckpt/"}, {"serialize_protocol": "flatbuffers"}):
        for device in ("cpu", "cuda_emu"):
            with pytest.raises(Exception):
                Processor(os.path.join(root, "v1"), dict(reference_config, **bad), device=device)
