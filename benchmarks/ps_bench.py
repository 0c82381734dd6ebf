#!/usr/bin/env python
"""Asynchronous parameter-server data plane: sparse pull + push throughput of the native TCP path (csrc/host/ps_server.cc) vs torch.distributed.rpc.

  python benchmarks/ps_bench.py --num_ps 2 --workers 2 --tables 8 --ids 8192 --dim 16 --steps 50

Every worker runs `steps` iterations of: ONE fused pull of `tables` x `ids` keys (power-law ids over 1 M per table), ONE asynchronous push of the
matching gradients.  Reports keys/s (pulled + pushed) summed over the workers, per transport.  CPU only."""
from __future__ import annotations

import argparse
import json
import os
import socket
import sys
import tempfile
import time

import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _role(rank, a, port, transport, tmp):
    os.environ["DEEPREC_PS_TRANSPORT"] = transport
    os.environ.setdefault("OMP_NUM_THREADS", "2"); os.environ.setdefault("DEEPREC_HOST_THREADS", "2")
    torch.set_num_threads(2)
    from deeprec_b200.parallel import ps
    if rank < a.num_ps:
        ps.run_ps(rank, a.num_ps, a.workers, port)
        return
    j = rank - a.num_ps
    client = ps.PSClient(j, a.num_ps, a.workers, port)
    embs = [client.create_embedding(f"t{t}", a.dim, optimizer="adagrad", lr=0.05, seed=t) for t in range(a.tables)]
    g = torch.Generator().manual_seed(100 + j)
    batches = [[(torch.rand(a.ids, generator=g) ** 3 * 1_000_000).long() for _ in range(a.tables)] for _ in range(8)]
    grads = [torch.randn(a.ids, a.dim, generator=g) * 0.01 for _ in range(a.tables)]

    def step(s):
        ids = batches[s % 8]
        client.pull_many([(e.name, i) for e, i in zip(embs, ids)])
        client.push_many([(e.name, i, gr) for e, i, gr in zip(embs, ids, grads)])
    for s in range(5):
        step(s)
    client.wait()
    t0 = time.perf_counter()
    for s in range(a.steps):
        step(s)
    client.wait()
    dt = time.perf_counter() - t0
    with open(os.path.join(tmp, f"w{j}.json"), "w") as f:
        json.dump({"seconds": dt, "native_ps": sorted(p for p, c in client._conns.items() if c is not None)}, f)
    client.shutdown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num_ps", type=int, default=2)
    ap.add_argument("--workers", type=int, default=2)
    ap.add_argument("--tables", type=int, default=8)
    ap.add_argument("--ids", type=int, default=8192)
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    lines = []
    for transport in ("rpc", "native"):
        tmp = tempfile.mkdtemp()
        port = _free_port()
        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_role, args=(r, a, port, transport, tmp)) for r in range(a.num_ps + a.workers)]
        [p.start() for p in procs]
        [p.join(600) for p in procs]
        ws = [json.load(open(os.path.join(tmp, f"w{j}.json"))) for j in range(a.workers)]
        keys = 2 * a.tables * a.ids * a.steps * a.workers                   # pulled + pushed
        sec = max(w["seconds"] for w in ws)
        rec = {"metric": "async PS sparse pull + push", "transport": transport, "num_ps": a.num_ps, "workers": a.workers, "tables": a.tables, "ids_per_table": a.ids,
               "dim": a.dim, "steps": a.steps, "keys_per_s": keys / sec, "ms_per_step": sec / a.steps * 1e3,
               "payload_MB_per_step_per_worker": a.tables * a.ids * (8 + 8 + 2 * a.dim * 4) / 1e6, "native_ps": ws[0]["native_ps"], "vcpus": os.cpu_count()}
        print(json.dumps(rec), flush=True)
        lines.append(json.dumps(rec))
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
