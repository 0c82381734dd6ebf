"""``python -m deeprec_b200.parallel.launch [--nproc N] script.py args...``: one process per visible GPU
(python/distribute/launch.py:34-319 in the reference, which synthesises TF_CONFIG/ports; here RANK/LOCAL_RANK/WORLD_SIZE +
a 127.0.0.1 rendezvous for torch.distributed)."""
from __future__ import annotations

import argparse
import os
import socket
import subprocess
import sys


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--nproc", type=int, default=0, help="processes (default: number of visible GPUs, or 1)")
    ap.add_argument("script")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    n = a.nproc
    if n <= 0:
        try:
            import torch
            n = max(1, torch.cuda.device_count())
        except Exception:
            n = 1
    port = _free_port()
    procs = []
    for r in range(n):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(n), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, a.script] + a.args, env=env))
    rc = 0
    for p in procs:
        rc = rc or p.wait()
    return rc


if __name__ == "__main__":
    sys.exit(main())
