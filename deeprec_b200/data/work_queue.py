"""WorkQueue: a global, elastic queue of work items (files / slices) that workers ``take()`` from.

Parity: python/ops/work_queue.py:113-598 + kernels/work_queue_ops.cc -- epochs, shuffle, resumable position
(save/restore), ``input_dataset()`` / ``input_producer()`` adapters, ``add_summary()`` progress.  The queue itself is
native (csrc/host/io_runtime.cc WorkQueue).  In multi-process runs rank 0 owns the queue and serves ``take`` over the
torch.distributed TCP store (the reference places it on the chief / PS); single-process use needs no store."""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Iterator, Optional

from .. import _native


class WorkQueue:
    def __init__(self, works: Iterable[str], num_epochs: int = 1, shuffle: bool = True, seed: int = 0, num_slices: Optional[int] = None,
                 name: str = "work_queue", store=None, rank: int = 0):
        works = [str(w) for w in works]
        if num_slices and num_slices > 1:          # slice each work item (work_queue.py: num_slices)
            works = [f"{w}?slice={i}/{num_slices}" for w in works for i in range(num_slices)]
        self.name, self.rank, self.store = name, rank, store
        self.lib = _native.host()
        self.h = None
        if store is None or rank == 0:
            self.h = self.lib.dr_wq_create("\n".join(works).encode(), int(num_epochs), int(shuffle), C.c_uint64(seed))
        self._taken = 0

    # ---- local (owner) operations -------------------------------------------------------------------------
    def _take_local(self) -> Optional[str]:
        buf = C.create_string_buffer(4096)
        n = self.lib.dr_wq_take(self.h, buf, 4096)
        return None if n < 0 else buf.value.decode()

    def take(self) -> Optional[str]:
        """Next work item or None when every epoch is consumed."""
        if self.store is None:
            w = self._take_local()
        elif self.rank == 0:
            lock = getattr(self, "_svc_lock", None)
            if lock is not None:
                with lock:
                    w = self._take_local()
            else:
                w = self._take_local()
        else:
            # remote take: a monotonically increasing ticket; rank 0 answers through the store (serve_once)
            t = self.store.add(f"{self.name}/ticket", 1)
            self.store.wait([f"{self.name}/ans/{t}"])
            v = self.store.get(f"{self.name}/ans/{t}").decode()
            w = None if v == "\0" else v
        if w is not None:
            self._taken += 1
        return w

    def serve_pending(self) -> int:
        """Rank 0: answer remote tickets issued so far (call from a service thread)."""
        served = 0
        store = getattr(self, "_svc_store", None) or self.store
        issued = int(store.add(f"{self.name}/ticket", 0))
        done = getattr(self, "_served", 0)
        for t in range(done + 1, issued + 1):
            w = self._take_local()
            store.set(f"{self.name}/ans/{t}", w if w is not None else "\0")
            served += 1
        self._served = issued
        return served

    def start_service(self, poll_s: float = 0.002) -> None:
        """Rank 0: answer remote ``take()`` calls from a daemon thread until ``stop_service()`` (the reference hosts the queue on
        the chief / a PS; here the owner is rank 0 and the transport is the torch.distributed store)."""
        import threading
        import time
        if self.store is None or self.rank != 0 or getattr(self, "_svc", None) is not None:
            return
        self._svc_stop = threading.Event()
        self._svc_lock = getattr(self, "_svc_lock", None) or threading.Lock()
        # a store client is one socket: a blocking wait() on the caller's thread would stall the service, so it gets its own
        self._svc_store = self._clone_store(self.store)

        def loop():
            while not self._svc_stop.is_set():
                with self._svc_lock:
                    n = self.serve_pending()
                if n == 0:
                    time.sleep(poll_s)

        self._svc = threading.Thread(target=loop, name=f"{self.name}-service", daemon=True)
        self._svc.start()

    @staticmethod
    def _clone_store(store):
        try:
            import torch.distributed as dist
            if isinstance(store, dist.TCPStore):
                return dist.TCPStore(store.host, store.port, None, False, timeout=store.timeout)
        except Exception:
            pass
        return store

    def stop_service(self) -> None:
        svc = getattr(self, "_svc", None)
        if svc is not None:
            self._svc_stop.set()
            svc.join()
            self._svc = None
            with self._svc_lock:
                self.serve_pending()          # tickets issued while stopping still get an answer

    def add(self, work: str) -> None:
        self.lib.dr_wq_add(self.h, str(work).encode())

    def remaining(self) -> int:
        return int(self.lib.dr_wq_remaining(self.h))

    # ---- resumable position (work_queue.py save/restore ops) -------------------------------------------------
    def state_dict(self) -> dict:
        e, p, t = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        self.lib.dr_wq_state(self.h, C.byref(e), C.byref(p), C.byref(t))
        return {"epoch": e.value, "pos": p.value, "taken": t.value}

    def load_state_dict(self, st: dict) -> None:
        self.lib.dr_wq_restore(self.h, int(st["epoch"]), int(st["pos"]))

    def add_summary(self) -> dict:
        return self.state_dict()

    # ---- adapters -------------------------------------------------------------------------------------------
    def input_producer(self) -> Iterator[str]:
        while True:
            w = self.take()
            if w is None:
                return
            yield w

    def input_dataset(self, reader) -> Iterator:
        """Chain ``reader(work)`` iterables over the taken work items (``WorkQueue.input_dataset``)."""
        for w in self.input_producer():
            yield from reader(w)

    def __del__(self):
        try:
            if self.h:
                self.lib.dr_wq_destroy(self.h)
        except Exception:
            pass
