"""``tf.stream(id)`` / multi-stream and ``tf.train.mark_target_node`` analogues (core/graph/stream_subgraph.{h,cc},
training/monitored_session.py:449 in the reference).

In the reference a graph pass assigns stream ids to sub-graphs; in eager PyTorch the user (or the engine) places work on a stream
directly.  ``stream(i)`` hands out one persistent side stream per (device, id) and takes care of the fork/join ordering with the
current stream, so ``with dr.utils.stream(1): emb = lookup(...)`` overlaps with whatever follows on the main stream.
"""
from __future__ import annotations

import contextlib
from typing import Dict, Iterable, Tuple

import torch

_STREAMS: Dict[Tuple[int, int], "torch.cuda.Stream"] = {}
_TARGETS: list = []


@contextlib.contextmanager
def stream(stream_id: int, device=None, join: bool = True):
    """Run the body on side stream ``stream_id`` (0 = current stream).  Fork: the side stream waits for work already queued on the
    current stream; join (default): the current stream waits for the body when the block exits."""
    if stream_id == 0 or not torch.cuda.is_available():
        yield None
        return
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    key = (dev.index or 0, int(stream_id))
    s = _STREAMS.get(key)
    if s is None:
        s = _STREAMS[key] = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    s.wait_stream(main)
    with torch.cuda.stream(s):
        yield s
    if join:
        main.wait_stream(s)


def mark_target_node(tensors: Iterable) -> list:
    """``tf.train.mark_target_node``: names the tensors at which SmartStage must cut when the automatic boundary (the data loader)
    is not the desired one.  Eager-mode meaning: the returned list is what ``smart_stage(..., targets=...)`` stages."""
    t = list(tensors) if not torch.is_tensor(tensors) else [tensors]
    _TARGETS.extend(t)
    return t


def target_nodes() -> list:
    return list(_TARGETS)
