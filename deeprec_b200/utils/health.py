"""Failure detection, fault injection and the pieces of elastic recovery that live outside the checkpoint code.

  * ``StepWatchdog``     a training / serving loop calls ``tick()``; if no tick arrives for ``timeout_s`` the watchdog dumps every
                         thread's Python stack (faulthandler) and runs a callback (default: hard-exit with code 86 so a supervisor
                         restarts the job, which then resumes from the last full + incremental checkpoints — Trainer does that).
                         A hung NVLink flag barrier (a dead peer never arrives) or a stuck data loader surfaces this way.
  * ``HeartbeatMonitor`` every rank publishes a monotonically increasing beat in a ``torch.distributed`` store (TCPStore / FileStore);
                         ``dead_ranks()`` lists peers whose beat has not advanced for ``timeout_s``.
  * ``FaultInjector``    ``DEEPREC_FAULT="step=<n>,kind=<exception|exit|hang>[,rank=<r>]"`` makes a chosen step fail, so recovery
                         paths are exercised in tests (the reference has no fault-injection tooling).
"""
from __future__ import annotations

import faulthandler
import os
import sys
import threading
import time
from typing import Callable, Dict, List, Optional


class StepWatchdog:
    EXIT_CODE = 86

    def __init__(self, timeout_s: float, on_stall: Optional[Callable[[float], None]] = None, poll_s: float = 0.5):
        self.timeout_s, self.poll_s = float(timeout_s), poll_s
        self.on_stall = on_stall or self._default_on_stall
        self._last = time.monotonic()
        self._stop = threading.Event()
        self._fired = False
        self.stalls = 0
        self._t = threading.Thread(target=self._loop, name="deeprec-watchdog", daemon=True)
        self._t.start()

    def tick(self) -> None:
        self._last = time.monotonic()
        self._fired = False

    def _default_on_stall(self, idle: float) -> None:
        sys.stderr.write(f"[deeprec watchdog] no training step for {idle:.1f}s -- dumping stacks and exiting {self.EXIT_CODE}\n")
        faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
        sys.stderr.flush()
        os._exit(self.EXIT_CODE)

    def _loop(self) -> None:
        while not self._stop.wait(self.poll_s):
            idle = time.monotonic() - self._last
            if idle > self.timeout_s and not self._fired:
                self._fired = True
                self.stalls += 1
                self.on_stall(idle)

    def close(self) -> None:
        self._stop.set()
        self._t.join(timeout=2 * self.poll_s + 1)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class HeartbeatMonitor:
    """Liveness of the ranks of a job through a shared ``torch.distributed`` store."""

    def __init__(self, store, rank: int, world_size: int, interval_s: float = 1.0, timeout_s: float = 10.0):
        self.store, self.rank, self.world, self.interval_s, self.timeout_s = store, rank, world_size, interval_s, timeout_s
        self._beat = 0
        self._seen: Dict[int, tuple] = {}          # rank -> (last beat value, local time it changed)
        self._stop = threading.Event()
        self._publish()
        self._t = threading.Thread(target=self._loop, name="deeprec-heartbeat", daemon=True)
        self._t.start()

    def _publish(self) -> None:
        self._beat += 1
        self.store.set(f"deeprec/hb/{self.rank}", str(self._beat))

    def _loop(self) -> None:
        while not self._stop.wait(self.interval_s):
            try:
                self._publish()
            except Exception:
                return                              # the store is gone: the job is shutting down

    def _poll(self) -> None:
        now = time.monotonic()
        for r in range(self.world):
            if r == self.rank:
                continue
            try:
                v = int(self.store.get(f"deeprec/hb/{r}"))
            except Exception:
                v = -1
            last = self._seen.get(r)
            if last is None or last[0] != v:
                self._seen[r] = (v, now)

    def dead_ranks(self) -> List[int]:
        self._poll()
        now = time.monotonic()
        return [r for r, (v, t) in self._seen.items() if now - t > self.timeout_s]

    def close(self) -> None:
        self._stop.set()
        self._t.join(timeout=self.interval_s + 1)


class InjectedFault(RuntimeError):
    pass


class FaultInjector:
    """Parse ``DEEPREC_FAULT`` once; ``maybe_fail(step)`` raises / exits / hangs at the configured step (on the configured rank)."""

    def __init__(self, spec: Optional[str] = None, rank: int = 0):
        spec = spec if spec is not None else os.environ.get("DEEPREC_FAULT", "")
        self.step, self.kind, self.rank = -1, "exception", None
        for part in filter(None, (p.strip() for p in spec.split(","))):
            k, _, v = part.partition("=")
            if k == "step":
                self.step = int(v)
            elif k == "kind":
                self.kind = v
            elif k == "rank":
                self.rank = int(v)
        self.my_rank = rank
        self.fired = False

    def maybe_fail(self, step: int) -> None:
        if self.step < 0 or self.fired or step != self.step or (self.rank is not None and self.rank != self.my_rank):
            return
        self.fired = True
        if self.kind == "exit":
            os._exit(13)
        if self.kind == "hang":
            while True:
                time.sleep(3600)
        raise InjectedFault(f"injected fault at step {step}")
