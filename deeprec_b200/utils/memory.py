"""Memory planning (reference: docs/docs_en/{CPU,GPU}-Memory-Optimization.md -- MemoryPlanner + TensorPoolAllocator,
``TF_GPU_ALLOCATOR=tensorpool``, ``START/STABLE/MAX_STATISTIC_STEP``).

The flagship engine needs none of this (static buffers + one CUDA graph per step: there is no per-step allocation).  Models
written against the generic framework API (model zoo, user modules) allocate activations every step; for those:

* ``HostTensorPool``  -- planned pool for host-side per-step buffers (staging / preprocessing tensors): the first
  ``collect_steps`` steps are observed, then every recurring buffer is served from a pre-carved slab in O(1).
* ``enable_gpu_tensorpool()`` -- installs the same planner over ``cudaMalloc`` as PyTorch's CUDA allocator
  (``CUDAPluggableAllocator`` -> ``dr_tp_cuda_malloc / dr_tp_cuda_free`` in libdeeprec_cuda.so); call ``gpu_tensorpool_step()`` once
  per training step.  Blocks are stream-tagged, so reuse never crosses streams.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Dict, Sequence

import numpy as np
import torch

from .. import _native
from .. import build as _build

_STAT_NAMES = ("phase", "steps", "pool_bytes", "pool_hits", "pool_misses", "small_bypass", "backend_allocs", "live_pool_blocks", "replans")
_BOUND = False


def _lib():
    global _BOUND
    L = _native.host()
    if not _BOUND:
        vp, i64 = C.c_void_p, C.c_int64
        L.dr_tp_create.restype, L.dr_tp_create.argtypes = vp, [i64, C.c_int, C.c_int]
        L.dr_tp_destroy.argtypes = [vp]
        L.dr_tp_set_start_step.argtypes = [vp, C.c_int]
        L.dr_tp_alloc.restype, L.dr_tp_alloc.argtypes = vp, [vp, i64, C.c_uint64]
        L.dr_tp_free.argtypes = [vp, vp]
        L.dr_tp_step_end.argtypes = [vp]
        L.dr_tp_stats.argtypes = [vp, C.POINTER(i64)]
        L.dr_tp_class_bytes.restype, L.dr_tp_class_bytes.argtypes = i64, [i64]
        _BOUND = True
    return L


class HostTensorPool:
    """Planned host allocator.  ``empty(shape, dtype)`` returns a tensor backed by pool memory; the block goes back to the pool when
    the tensor (and every view of it) is garbage collected."""

    def __init__(self, small_threshold: int = 4096, collect_steps: int = 3, replan_misses: int = 8, start_step: int = 0):
        self.L = _lib()
        self.h = self.L.dr_tp_create(small_threshold, collect_steps, replan_misses)
        if start_step > 0:
            self.L.dr_tp_set_start_step(self.h, start_step)       # steps before it (initialisation) are not part of the plan

    @classmethod
    def from_env(cls, **kw) -> "HostTensorPool":
        """The reference's switches (CPU-Memory-Optimization.md): ``START_STATISTIC_STEP`` (first step whose allocations count),
        ``STABLE_STATISTIC_STEP`` / ``MAX_STATISTIC_STEP`` (length of the collection window; the smaller one wins -- re-planning on
        misses replaces the reference's stability test), ``ENABLE_MEMORY_OPTIMIZATION=0`` (every request goes to malloc)."""
        env = os.environ
        if env.get("ENABLE_MEMORY_OPTIMIZATION", "1").strip() in ("0", "false", "False"):
            kw.setdefault("small_threshold", 1 << 62)
        window = [int(env[k]) for k in ("STABLE_STATISTIC_STEP", "MAX_STATISTIC_STEP") if env.get(k, "").isdigit() and int(env[k]) > 0]
        if window:
            kw.setdefault("collect_steps", min(window))
        if env.get("START_STATISTIC_STEP", "").isdigit():
            kw.setdefault("start_step", int(env["START_STATISTIC_STEP"]))
        return cls(**kw)

    def empty(self, shape: Sequence[int], dtype: torch.dtype = torch.float32) -> torch.Tensor:
        shape = tuple(int(s) for s in shape)
        nbytes = int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=dtype).element_size()
        if nbytes == 0:
            return torch.empty(shape, dtype=dtype)
        p = self.L.dr_tp_alloc(self.h, nbytes, 0)
        if not p:
            raise MemoryError(f"HostTensorPool: cannot allocate {nbytes} bytes")
        buf = (C.c_uint8 * nbytes).from_address(p)
        weakref.finalize(buf, self._release, self.L, self.h, p)          # fires when the last tensor view drops the buffer
        return torch.frombuffer(buf, dtype=dtype).view(shape)

    @staticmethod
    def _release(L, h, p):
        if h:
            L.dr_tp_free(h, p)

    def step_end(self) -> None:
        self.L.dr_tp_step_end(self.h)

    def stats(self) -> Dict[str, int]:
        out = (C.c_int64 * 9)()
        self.L.dr_tp_stats(self.h, out)
        return dict(zip(_STAT_NAMES, (int(v) for v in out)))

    def class_bytes(self, nbytes: int) -> int:
        return int(self.L.dr_tp_class_bytes(nbytes))


_GPU = None


def enable_gpu_tensorpool() -> None:
    """Make the planned pool PyTorch's CUDA allocator.  Must run before the first CUDA allocation of the process."""
    global _GPU
    if _GPU is not None:
        return
    path = os.path.join(_build.LIB, "libdeeprec_cuda.so")
    if not os.path.exists(path):
        path = _build.build_cuda()
    alloc = torch.cuda.memory.CUDAPluggableAllocator(path, "dr_tp_cuda_malloc", "dr_tp_cuda_free")
    torch.cuda.memory.change_current_allocator(alloc)
    lib = C.CDLL(path)
    lib.dr_tp_cuda_step_end.argtypes = [C.c_int]
    lib.dr_tp_cuda_stats.argtypes = [C.c_int, C.POINTER(C.c_int64)]
    _GPU = (alloc, lib)


def gpu_tensorpool_step(device: int | None = None) -> None:
    if _GPU is None:
        raise RuntimeError("enable_gpu_tensorpool() was not called")
    _GPU[1].dr_tp_cuda_step_end(torch.cuda.current_device() if device is None else device)


def gpu_tensorpool_stats(device: int | None = None) -> Dict[str, int]:
    if _GPU is None:
        raise RuntimeError("enable_gpu_tensorpool() was not called")
    out = (C.c_int64 * 9)()
    _GPU[1].dr_tp_cuda_stats(torch.cuda.current_device() if device is None else device, out)
    return dict(zip(_STAT_NAMES, (int(v) for v in out)))


def enable_cuda_malloc_async() -> bool:
    """Stream-ordered driver allocator (``cudaMallocAsync`` memory pools) as PyTorch's CUDA allocator -- the reference's
    ``TF_GPU_ALLOCATOR=cuda_malloc_async`` (``common_runtime/gpu/gpu_cudamallocasync_allocator.cc``).  Must run before the first CUDA
    allocation of the process (PyTorch fixes its allocator backend at that point); returns False when it is too late."""
    if torch.cuda.is_initialized():
        return torch.cuda.get_allocator_backend() == "cudaMallocAsync"
    conf = [c for c in os.environ.get("PYTORCH_CUDA_ALLOC_CONF", "").split(",") if c and not c.startswith("backend:")]
    os.environ["PYTORCH_CUDA_ALLOC_CONF"] = ",".join(conf + ["backend:cudaMallocAsync"])
    setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or getattr(torch.cuda.memory, "_set_allocator_settings", None)
    try:
        if setter is not None:
            setter(os.environ["PYTORCH_CUDA_ALLOC_CONF"])
    except Exception:
        pass                                     # builds that read the variable at first use only
    return True


def maybe_enable_from_env() -> bool:
    """``TF_GPU_ALLOCATOR=tensorpool | cuda_malloc_async`` / ``DEEPREC_GPU_ALLOCATOR=...`` (the reference's switch)."""
    v = os.environ.get("DEEPREC_GPU_ALLOCATOR", os.environ.get("TF_GPU_ALLOCATOR", "")).strip().lower()
    if v == "tensorpool" and torch.cuda.is_available():
        enable_gpu_tensorpool()
        return True
    if v in ("cuda_malloc_async", "cudamallocasync"):
        return enable_cuda_malloc_async()
    return False
