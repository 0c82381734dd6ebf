"""SmartStage / tf.staged: asynchronous input pipeline with bounded staging buffers.

Reference behaviour: ``tf.staged`` marks a cut; ``TensorBufferPut/Take/Cancel/Close/Size`` form a bounded
multi-producer buffer (python/ops/prefetch.py:43,92-386, kernels/tensor_buffer_ops.cc:94,124); ``PrefetchRunner`` threads
keep running the producer sub-graph (cc/training/prefetch_runner.cc); SmartStage finds the cut automatically at the
IO/preprocess boundary (core/graph/smart_stage_pass.cc:193-419), ``stage_subgraph_on_cpu`` keeps it off the GPU stream and
``TensorPackTransH2D`` packs the staged tensors into ONE pinned buffer / ONE H2D copy (gpu_stage_pack_trans_pass.cc).

Here (eager PyTorch, no graph pass needed): the cut is by construction at the producer callable / iterable boundary;
producer threads run the CPU side (read + parse + hash/unique), pack all tensors of a batch into one pinned host block
and enqueue a ticket in the native bounded queue (csrc/host/io_runtime.cc StagingQueue); the consumer side issues ONE
``cudaMemcpyAsync`` per batch on a dedicated copy stream, ``capacity`` batches ahead of compute.
"""
from __future__ import annotations

import ctypes as C
import itertools
import threading
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import torch

from .. import _native

Batch = Union[torch.Tensor, Sequence[torch.Tensor], Dict[str, torch.Tensor]]


class StagingBuffer:
    """Bounded buffer of staged items (TensorBuffer analogue).  Payloads stay in Python; blocking / capacity /
    timeout / cancel / close semantics are the native queue's."""

    def __init__(self, capacity: int = 1, timeout_millis: int = -1):
        self.lib = _native.host()
        self.h = self.lib.dr_stage_create(int(capacity))
        self.timeout = int(timeout_millis)
        self._items: Dict[int, Any] = {}
        self._ids = itertools.count(1)
        self._lock = threading.Lock()

    def put(self, item: Any, timeout_millis: Optional[int] = None) -> bool:
        t = next(self._ids)
        with self._lock:
            self._items[t] = item
        rc = self.lib.dr_stage_put(self.h, t, self.timeout if timeout_millis is None else int(timeout_millis))
        if rc != 0:
            with self._lock:
                self._items.pop(t, None)
            if rc == 1:
                raise TimeoutError("StagingBuffer.put timed out")
            return False           # closed / cancelled
        return True

    def take(self, timeout_millis: Optional[int] = None) -> Any:
        t = C.c_int64(0)
        rc = self.lib.dr_stage_take(self.h, C.byref(t), self.timeout if timeout_millis is None else int(timeout_millis))
        if rc == 1:
            raise TimeoutError("StagingBuffer.take timed out")
        if rc == 2:
            raise StopIteration
        with self._lock:
            return self._items.pop(t.value)

    def size(self) -> int:
        return int(self.lib.dr_stage_size(self.h))

    def cancel(self) -> int:
        """Drop everything staged and reject producers until ``resume`` (TensorBufferCancel)."""
        buf = (C.c_int64 * 4096)()
        n = int(self.lib.dr_stage_cancel(self.h, buf, 4096))
        with self._lock:
            for i in range(min(n, 4096)):
                self._items.pop(buf[i], None)
        return n

    def resume(self) -> None:
        self.lib.dr_stage_resume(self.h)

    def close(self) -> None:
        self.lib.dr_stage_close(self.h)

    def __del__(self):
        try:
            self.lib.dr_stage_close(self.h)
            self.lib.dr_stage_destroy(self.h)
        except Exception:
            pass


class _PinnedPool:
    """Recycled pinned host blocks: ``cudaHostAlloc`` costs milliseconds, a training step costs one -- a block is pinned once
    and handed back by the consumer when the H2D copy that read it has completed (event query, no host sync)."""

    def __init__(self):
        self._free: Dict[int, List[torch.Tensor]] = {}
        self._pending: List[Tuple[Any, torch.Tensor]] = []
        self._lock = threading.Lock()

    def take(self, nbytes: int) -> torch.Tensor:
        cap = 1 << max(12, (max(1, nbytes) - 1).bit_length())
        with self._lock:
            self._reap()
            lst = self._free.get(cap)
            if lst:
                return lst.pop()
        return torch.empty(cap, dtype=torch.uint8).pin_memory()

    def give_back(self, block: torch.Tensor, event) -> None:
        with self._lock:
            self._pending.append((event, block))
            self._reap()

    def _reap(self) -> None:
        keep = []
        for ev, blk in self._pending:
            if ev is None or ev.query():
                self._free.setdefault(blk.numel(), []).append(blk)
            else:
                keep.append((ev, blk))
        self._pending = keep


class PackedHostBatch:
    """All tensors of a batch packed back-to-back (256 B aligned) in ONE pinned host block (TensorPackTransH2D,
    gpu_stage_pack_trans_pass.cc).  Tensors that are ALREADY pinned are not repacked: they are copied straight from where
    they are (zero host copies), into the same packed device layout."""

    def __init__(self, tensors: List[torch.Tensor], pin: bool = True, pool: Optional[_PinnedPool] = None):
        self.meta = []
        off = 0
        for t in tensors:
            nb = t.numel() * t.element_size()
            self.meta.append((off, nb, t.dtype, tuple(t.shape)))
            off += (nb + 255) // 256 * 256
        self.nbytes = off
        self.pool = pool
        self.sources: Optional[List[torch.Tensor]] = None
        cuda = pin and torch.cuda.is_available()
        if cuda and all(t.is_pinned() and t.is_contiguous() for t in tensors):
            self.sources, self.block = list(tensors), None
            return
        if cuda and pool is not None:
            self.block = pool.take(max(off, 1))
        else:
            self.block = torch.empty(max(off, 1), dtype=torch.uint8)
            if cuda:
                self.block = self.block.pin_memory()
        for (o, nb, _, _), t in zip(self.meta, tensors):
            if nb:
                self.block[o:o + nb].copy_(t.contiguous().view(-1).view(torch.uint8))

    def copy_into(self, dev_block: torch.Tensor) -> List[torch.Tensor]:
        """Async H2D into a caller-owned device block (current stream); returns device views."""
        if self.sources is not None:
            for (o, nb, _, _), t in zip(self.meta, self.sources):
                if nb:
                    dev_block[o:o + nb].copy_(t.view(-1).view(torch.uint8), non_blocking=True)
        else:
            dev_block[: self.nbytes].copy_(self.block[: self.nbytes], non_blocking=True)      # ONE H2D copy
        return [dev_block[o:o + nb].view(dt).view(shape) for (o, nb, dt, shape) in self.meta]

    def to_device(self, device, stream=None) -> List[torch.Tensor]:
        """H2D into a fresh device block; returns device views."""
        ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
        with ctx:
            d = torch.empty(max(self.nbytes, 1), dtype=torch.uint8, device=device)
            return self.copy_into(d)

    def release(self, event=None) -> None:
        if self.pool is not None and self.block is not None:
            self.pool.give_back(self.block, event)
            self.block = None

    def unpack_host(self) -> List[torch.Tensor]:
        if self.sources is not None:
            return list(self.sources)
        return [self.block[o:o + nb].view(dt).view(shape) for (o, nb, dt, shape) in self.meta]


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _flatten(b: Batch) -> Tuple[List[torch.Tensor], Callable[[List[torch.Tensor]], Batch]]:
    if torch.is_tensor(b):
        return [b], lambda ts: ts[0]
    if isinstance(b, dict):
        keys = list(b.keys())
        return [b[k] for k in keys], lambda ts: dict(zip(keys, ts))
    typ = type(b)
    return list(b), lambda ts: typ(ts) if typ in (list, tuple) else tuple(ts)


class PrefetchRunner:
    """Producer threads that keep a StagingBuffer full (cc/training/prefetch_runner.cc)."""

    def __init__(self, buffer: StagingBuffer, produce: Callable[[], Any], num_threads: int = 1, name: str = "prefetch"):
        self.buffer, self.produce, self.n, self.name = buffer, produce, max(1, num_threads), name
        self.threads: List[threading.Thread] = []
        self._stop = threading.Event()
        self.error: Optional[BaseException] = None
        self._live = self.n
        self._live_lock = threading.Lock()

    def start(self) -> "PrefetchRunner":
        for i in range(self.n):
            t = threading.Thread(target=self._loop, name=f"{self.name}-{i}", daemon=True)
            t.start()
            self.threads.append(t)
        return self

    def _loop(self) -> None:
        try:
            while not self._stop.is_set():
                item = self.produce()
                if not self.buffer.put(item):
                    break
        except StopIteration:
            pass
        except BaseException as e:          # surfaced to the consumer
            self.error = e
        finally:
            with self._live_lock:
                self._live -= 1
                last = self._live == 0
            if last or self.error is not None:
                self.buffer.close()         # only the last producer closes: in-flight items of its peers are not lost

    def stop(self) -> None:
        self._stop.set()
        self.buffer.close()
        for t in self.threads:
            t.join(timeout=5)


class Staged:
    """Iterator over staged batches.  ``device`` given => every batch arrives on the device via one packed H2D copy
    issued on a side stream ``capacity`` batches ahead; the consumer's stream waits on the copy event only."""

    def __init__(self, source: Union[Iterable[Batch], Callable[[], Batch]], capacity: int = 2, num_threads: int = 1,
                 timeout_millis: int = -1, device: Optional[torch.device] = None, pin: bool = True,
                 preprocess: Optional[Callable[[Batch], Batch]] = None, name: str = "staged"):
        self.device = torch.device(device) if device is not None else None
        self.preprocess = preprocess
        self.pin = pin and self.device is not None and self.device.type == "cuda"
        self.buffer = StagingBuffer(capacity, timeout_millis)
        if callable(source):
            produce_raw = source
        else:
            it = iter(source)
            lock = threading.Lock()

            def produce_raw():
                with lock:
                    return next(it)
        self._produce_raw = produce_raw
        self._tls = threading.local()
        self._pool = _PinnedPool() if self.pin else None
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.pin else None
        self._ahead: List[Tuple[Any, Any, Any, int]] = []
        self._depth = max(1, capacity)
        # device-side ring of packed blocks: slot reuse is ordered by events, never by the caching allocator
        self._nslots = self._depth + 2
        self._ring: List[Optional[torch.Tensor]] = [None] * self._nslots
        self._ring_free: List[Any] = [None] * self._nslots
        self._next_slot = 0
        self._last_slot = -1
        self.runner = PrefetchRunner(self.buffer, self._produce, num_threads, name).start()

    def _produce(self):
        if self.pin and getattr(self._tls, "dev_set", False) is False:
            # CUDA's current device is per host thread and starts at 0: without this, the first pinned-memory call of a producer
            # thread of rank r > 0 creates a context on GPU 0 (hundreds of ms, and memory on the wrong device)
            torch.cuda.set_device(self.device)
            self._tls.dev_set = True
        b = self._produce_raw()
        if self.preprocess is not None:
            b = self.preprocess(b)             # CPU side of the cut (stage_subgraph_on_cpu)
        ts, rebuild = _flatten(b)
        if self.pin:
            return PackedHostBatch(ts, True, self._pool), rebuild
        return ts, rebuild

    def _issue(self) -> bool:
        try:
            item, rebuild = self.buffer.take()
        except StopIteration:
            if self.runner.error is not None:
                raise self.runner.error
            return False
        if self.pin:
            slot = self._next_slot
            self._next_slot = (slot + 1) % self._nslots
            blk = self._ring[slot]
            if blk is None or blk.numel() < item.nbytes:
                blk = self._ring[slot] = torch.empty(max(item.nbytes, 1), dtype=torch.uint8, device=self.device)
            with torch.cuda.stream(self.copy_stream):
                if self._ring_free[slot] is not None:
                    self.copy_stream.wait_event(self._ring_free[slot])     # the consumer finished with this slot's previous batch
                dts = item.copy_into(blk)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            item.release(ev)
            self._ahead.append((rebuild(dts), ev, item, slot))
        else:
            ts = item if self.device is None else [t.to(self.device) for t in item]
            self._ahead.append((rebuild(ts), None, item, -1))
        return True

    def __iter__(self) -> Iterator[Batch]:
        return self

    def __next__(self) -> Batch:
        while len(self._ahead) < self._depth and self._issue():
            pass
        if not self._ahead:
            raise StopIteration
        batch, ev, _keep, slot = self._ahead.pop(0)
        if ev is not None:
            cur = torch.cuda.current_stream(self.device)
            if self._last_slot >= 0:         # asking for the next batch == done with the previous one (on the consumer's stream)
                done = torch.cuda.Event()
                done.record(cur)
                self._ring_free[self._last_slot] = done
            self._last_slot = slot
            cur.wait_event(ev)
        return batch

    get = __next__

    def size(self) -> int:
        return self.buffer.size()

    def cancel(self) -> None:
        self.buffer.cancel()
        self._ahead.clear()

    def close(self) -> None:
        self.runner.stop()


def staged(features: Union[Iterable[Batch], Callable[[], Batch]], capacity: int = 1, num_threads: int = 1, timeout_millis: int = 300000,
           device=None, closed_exception_types=None, ignored_exception_types=None, use_stage_subgraph_thread_pool: bool = False,
           stage_subgraph_stream_id: int = 0, preprocess=None, name: str = "staged") -> Staged:
    """``tf.staged`` (python/ops/prefetch.py:92)."""
    return Staged(features, capacity, num_threads, timeout_millis, device, True, preprocess, name)


class SmartStageOptions:
    """``tf.SmartStageOptions`` / ConfigProto.do_smart_stage (+ stage_subgraph_on_cpu) analogue."""

    def __init__(self, capacity: int = 2, num_threads: int = 1, stage_subgraph_on_cpu: bool = True, timeout_millis: int = 300000):
        self.capacity, self.num_threads, self.stage_subgraph_on_cpu, self.timeout_millis = capacity, num_threads, stage_subgraph_on_cpu, timeout_millis


def smart_stage(dataset: Iterable[Batch], device, options: Optional[SmartStageOptions] = None,
                preprocess: Optional[Callable[[Batch], Batch]] = None) -> Staged:
    """Automatic staging of everything upstream of the model: dataset iteration + ``preprocess`` run on producer threads
    (CPU), packed pinned H2D, consumer only sees device tensors."""
    o = options or SmartStageOptions()
    return Staged(dataset, o.capacity, o.num_threads, o.timeout_millis, device, True, preprocess, "smart_stage")


def make_prefetch_hook(*stages: Staged):
    """``tf.make_prefetch_hook``: in eager mode the runners are started at construction; the hook only closes them."""
    class _Hook:
        def end(self):
            for s in stages:
                s.close()
    return _Hook()


class AsyncEmbeddingStage:
    """Async embedding lookup (python/training/async_embedding_stage.py): a second stage that runs the embedding lookup of
    batch N+1 while the dense net trains on batch N (stale by one step, like the reference's ``capacity``)."""

    def __init__(self, batches: Iterable[Batch], lookup_fn: Callable[[Batch], Any], device=None):
        self.it = iter(batches)
        self.lookup_fn = lookup_fn
        self.device = device
        self.stream = torch.cuda.Stream(device=device) if device is not None and torch.device(device).type == "cuda" else None
        self._next = None

    def _start(self):
        try:
            b = next(self.it)
        except StopIteration:
            self._next = None
            return
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                e = self.lookup_fn(b)
            ev = torch.cuda.Event(); ev.record(self.stream)
        else:
            e, ev = self.lookup_fn(b), None
        self._next = (b, e, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            self._start()
        if self._next is None:
            raise StopIteration
        b, e, ev = self._next
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        self._start()
        return b, e
