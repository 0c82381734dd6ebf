"""Model server: ``python -m deeprec_b200.serving.serve --model ctr=/models/ctr/v7 [--model cvr=...] --grpc_port 8501 --http_port 8500``.

Loads one native Processor per ``--model name=saved_model_dir`` (GPU runtime when CUDA is present, the CPU runtime otherwise or with
``--device cpu``) and puts the gRPC ``PredictService`` and / or the HTTP front-end in front of them.  ``--config`` is the ModelConfig JSON
(file or inline) every processor is initialised with (session_num, checkpoint_dir for hot updates, feature store, ...).  The bound ports
are written to ``--port_file`` as JSON once the servers listen (port 0 = pick a free one).
"""
import argparse
import json
import os
import signal
import threading


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--model", action="append", required=True, metavar="NAME=DIR")
    ap.add_argument("--config", default="{}", help="ModelConfig JSON: a file path or an inline JSON object")
    ap.add_argument("--device", default=None, choices=[None, "cuda", "cpu"])
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--grpc_port", type=int, default=-1, help="-1 = no gRPC endpoint, 0 = any free port")
    ap.add_argument("--http_port", type=int, default=-1, help="-1 = no HTTP endpoint")
    ap.add_argument("--grpc_workers", type=int, default=8)
    ap.add_argument("--port_file", default="")
    a = ap.parse_args(argv)
    if a.grpc_port < 0 and a.http_port < 0:
        ap.error("enable at least one of --grpc_port / --http_port")
    from .processor import Processor
    cfg = json.load(open(a.config)) if os.path.exists(a.config) else json.loads(a.config)
    procs = {}
    for spec in a.model:
        name, _, path = spec.partition("=")
        if not path:
            ap.error(f"--model expects NAME=DIR, got {spec!r}")
        procs[name] = Processor(path, dict(cfg), device=a.device)
    ports, grpc_server, stop = {}, None, threading.Event()
    if a.grpc_port >= 0:
        from .grpc_server import create_server
        grpc_server, ports["grpc"] = create_server(procs, f"{a.host}:{a.grpc_port}", max_workers=a.grpc_workers)
    http = None
    if a.http_port >= 0:
        import socket

        import uvicorn

        from .http_server import ServingBackend, create_app
        sock = socket.socket(); sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1); sock.bind((a.host, a.http_port)); sock.listen(128)
        ports["http"] = sock.getsockname()[1]
        http = uvicorn.Server(uvicorn.Config(create_app({n: ServingBackend.from_processor(p) for n, p in procs.items()}), log_level="warning"))
        threading.Thread(target=lambda: http.run(sockets=[sock]), daemon=True).start()
    signal.signal(signal.SIGTERM, lambda *_: stop.set())
    signal.signal(signal.SIGINT, lambda *_: stop.set())
    if a.port_file:
        tmp = a.port_file + ".tmp"
        with open(tmp, "w") as f:
            json.dump(ports, f)
        os.replace(tmp, a.port_file)
    print(f"serving {sorted(procs)} on {ports}", flush=True)
    stop.wait()
    if grpc_server is not None:
        grpc_server.stop(1.0).wait()
    if http is not None:
        http.should_exit = True
    for p in procs.values():
        p.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
