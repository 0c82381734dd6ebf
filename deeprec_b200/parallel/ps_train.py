"""Parameter-server training job launcher (the modelzoo's ``--protocol grpc | grpc++ | star_server`` + ``TF_CONFIG`` roles):

    # every role by hand (one process each, any host that can reach the master):
    python -m deeprec_b200.parallel.ps_train --job_name ps     --task_index 0 --num_ps 2 --num_workers 2 --master_port 29700
    python -m deeprec_b200.parallel.ps_train --job_name worker --task_index 1 --num_ps 2 --num_workers 2 --master_port 29700 --steps 200
    # or everything on this machine:
    python -m deeprec_b200.parallel.ps_train --spawn --num_ps 2 --num_workers 2 --steps 50

The model is a Wide&Deep-shaped network whose 26 categorical tables live on the parameter servers (``PSEmbedding``: pulls on the
forward, asynchronous pushes on the backward, one fused RPC per server and step) and whose dense parameters are PS-hosted too.
``--spare_ps k`` starts k extra, initially idle servers; ``--scale_at step:n`` re-shards onto n servers mid-run (elastic training).
"""
from __future__ import annotations

import argparse
import json
import socket
import sys
import time
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class PSWideDeep(nn.Module):
    def __init__(self, client, num_dense: int = 13, num_sparse: int = 26, dim: int = 16, hidden=(256, 128), optimizer: str = "adagrad", lr: float = 0.05):
        super().__init__()
        self.client = client
        self.embs = [client.create_embedding(f"C{i + 1}", dim, optimizer=optimizer, lr=lr, seed=i) for i in range(num_sparse)]
        layers, k = [], num_dense + num_sparse * dim
        for n in hidden:
            layers += [nn.Linear(k, n), nn.ReLU()]
            k = n
        self.deep = nn.Sequential(*layers, nn.Linear(k, 1))
        self.wide = nn.Linear(num_dense, 1)

    def forward(self, dense: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        from . import ps
        rows = ps.group_pull(self.client, self.embs, [ids[i] for i in range(len(self.embs))])          # ONE rpc per PS for all tables
        x = torch.cat([dense] + rows, dim=1)
        return (self.deep(x) + self.wide(dense)).squeeze(-1)


def run_worker(a) -> dict:
    from ..data import criteo_batch
    from . import ps
    total_ps = a.num_ps + a.spare_ps
    client = ps.PSClient(a.task_index, total_ps, a.num_workers, a.master_port, active_ps=a.num_ps)
    torch.manual_seed(0)
    model = PSWideDeep(client, optimizer=a.optimizer, lr=a.learning_rate)
    client.register_dense(model, lr=a.learning_rate)
    scale_step, scale_to = (int(x) for x in a.scale_at.split(":")) if a.scale_at else (-1, 0)
    cards = [1000] * 26
    losses: List[float] = []
    t0 = time.time()
    for step in range(a.steps):
        if step == scale_step and a.task_index == 0:
            moved = client.scale(scale_to)
            print(f"[worker0] step {step}: re-sharded onto {scale_to} servers, {moved} rows moved", flush=True)
        client.pull_dense()
        d, ids, y = criteo_batch(a.batch_size, 13, cards, seed=step * a.num_workers + a.task_index)
        loss = F.binary_cross_entropy_with_logits(model(d, ids), y)
        model.zero_grad()
        loss.backward()
        ps.push_gradients(client, model.embs)                 # asynchronous: no barrier between workers
        losses.append(float(loss.detach()))
        if a.log_every and step % a.log_every == 0:
            print(f"[worker{a.task_index}] global_step {step} loss {losses[-1]:.5f}", flush=True)
    client.wait()
    out = {"worker": a.task_index, "first_loss": losses[0], "last_loss": sum(losses[-5:]) / len(losses[-5:]),
           "samples_per_s": a.steps * a.batch_size / (time.time() - t0), "active_ps": client.num_ps}
    if a.result:
        with open(f"{a.result}.worker{a.task_index}.json", "w") as f:
            json.dump(out, f)
    client.shutdown()
    return out


def run_role(a) -> int:
    from . import ps
    if a.job_name == "ps":
        ps.run_ps(a.task_index, a.num_ps + a.spare_ps, a.num_workers, a.master_port, active_ps=a.num_ps)
    else:
        print(json.dumps(run_worker(a)), flush=True)
    return 0


def _spawn_entry(argv: List[str]) -> None:
    sys.exit(main(argv))


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--job_name", choices=["ps", "worker"], default="worker")
    p.add_argument("--task_index", type=int, default=0)
    p.add_argument("--num_ps", type=int, default=1)
    p.add_argument("--spare_ps", type=int, default=0)
    p.add_argument("--num_workers", type=int, default=1)
    p.add_argument("--master_port", type=int, default=0)
    p.add_argument("--steps", type=int, default=50)
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--learning_rate", type=float, default=0.05)
    p.add_argument("--optimizer", default="adagrad")
    p.add_argument("--scale_at", default="", help="step:new_num_ps -- elastic re-shard (worker 0 drives it)")
    p.add_argument("--log_every", type=int, default=10)
    p.add_argument("--result", default="", help="prefix of per-worker JSON result files")
    p.add_argument("--spawn", action="store_true", help="start every role on this machine")
    a = p.parse_args(argv)
    if not a.spawn:
        return run_role(a)
    import torch.multiprocessing as mp
    if not a.master_port:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            a.master_port = s.getsockname()[1]
    base = [f"--num_ps={a.num_ps}", f"--spare_ps={a.spare_ps}", f"--num_workers={a.num_workers}", f"--master_port={a.master_port}", f"--steps={a.steps}",
            f"--batch_size={a.batch_size}", f"--learning_rate={a.learning_rate}", f"--optimizer={a.optimizer}", f"--scale_at={a.scale_at}",
            f"--log_every={a.log_every}", f"--result={a.result}"]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_spawn_entry, args=(base + ["--job_name=ps", f"--task_index={i}"],)) for i in range(a.num_ps + a.spare_ps)]
    procs += [ctx.Process(target=_spawn_entry, args=(base + ["--job_name=worker", f"--task_index={j}"],)) for j in range(a.num_workers)]
    for pr in procs:
        pr.start()
    rc = 0
    for pr in procs:
        pr.join()
        rc = rc or (pr.exitcode or 0)
    return rc


if __name__ == "__main__":
    sys.exit(main())
