"""SessionGroup for arbitrary PyTorch models (direct_session_group.{h,cc} analogue): N sessions = N worker contexts, each
with its own CUDA stream and lock, sharing ONE set of read-only parameters / embedding tables.  ``run`` picks a session
round-robin or by ``hint % n`` (direct_session_group.h:72-89).  (The native ``Processor`` additionally owns per-session pinned
IO buffers and replays CUDA graphs; this class is the generic-module counterpart used for zoo models and tests.)"""
from __future__ import annotations

import itertools
import threading
from typing import Any, Callable, List, Optional

import torch


class _Session:
    def __init__(self, idx: int, device: Optional[torch.device]):
        self.idx, self.lock = idx, threading.Lock()
        self.stream = torch.cuda.Stream(device=device) if device is not None and device.type == "cuda" else None


class SessionGroup:
    def __init__(self, model: torch.nn.Module, session_num: int = 2, select_session_policy: str = "RR", device=None):
        self.model = model.eval()
        self.device = torch.device(device) if device is not None else None
        self.sessions: List[_Session] = [_Session(i, self.device) for i in range(max(1, session_num))]
        self.policy = select_session_policy.upper()
        self._rr = itertools.count()
        self._swap_lock = threading.Lock()

    def _pick(self, hint: Optional[int]) -> _Session:
        if self.policy == "MOD":
            h = hint if hint is not None else threading.get_ident()
            return self.sessions[h % len(self.sessions)]
        return self.sessions[next(self._rr) % len(self.sessions)]

    @torch.no_grad()
    def run(self, *inputs, session_id: Optional[int] = None, fn: Optional[Callable] = None) -> Any:
        s = self._pick(session_id)
        model = self.model
        with s.lock:
            if s.stream is not None:
                with torch.cuda.stream(s.stream):
                    out = (fn or model)(*inputs)
                s.stream.synchronize()
                return out
            return (fn or model)(*inputs)

    def swap_model(self, new_model: torch.nn.Module, warmup_inputs=None) -> None:
        """Full model update: warm the new model up, then swap atomically (requests in flight finish on the old one)."""
        new_model = new_model.eval()
        if warmup_inputs is not None:
            with torch.no_grad():
                new_model(*warmup_inputs)
        with self._swap_lock:
            self.model = new_model
