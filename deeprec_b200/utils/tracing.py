"""Tracing / profiling helpers: NVTX ranges + CUDA-event step timeline written as a chrome trace
(``--timeline N`` / ProfilerHook analogue, modelzoo/dlrm/train.py:521-524)."""
from __future__ import annotations

import contextlib
import json
import time
from typing import List

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class Timeline:
    """Collects (name, start, duration) spans; device spans via CUDA events, host spans via perf_counter."""

    def __init__(self):
        self.spans: List[dict] = []
        self._t0 = time.perf_counter()
        self._pending = []

    @contextlib.contextmanager
    def span(self, name: str, device: bool = False, tid: int = 0):
        if device and torch.cuda.is_available():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            host = time.perf_counter()
            e0.record()
            try:
                yield
            finally:
                e1.record()
                self._pending.append((name, host, e0, e1, tid))
        else:
            t = time.perf_counter()
            try:
                yield
            finally:
                self.spans.append({"name": name, "ph": "X", "ts": (t - self._t0) * 1e6, "dur": (time.perf_counter() - t) * 1e6, "pid": 0, "tid": tid})

    def flush(self) -> None:
        if self._pending:
            torch.cuda.synchronize()
            for name, host, e0, e1, tid in self._pending:
                self.spans.append({"name": name, "ph": "X", "ts": (host - self._t0) * 1e6, "dur": e0.elapsed_time(e1) * 1e3, "pid": 1, "tid": tid})
            self._pending.clear()

    def save(self, path: str) -> None:
        self.flush()
        with open(path, "w") as f:
            json.dump({"traceEvents": self.spans}, f)
