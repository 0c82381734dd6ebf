"""Bind a rank's host threads to the CPUs next to its GPU.

On an 8-GPU HGX box GPUs 0-3 hang off socket 0 and GPUs 4-7 off socket 1 (``nvidia-smi topo -m``: ``SYS`` = across the SMP interconnect).  A
rank whose input pipeline allocates its pinned staging buffers on the far socket pays the inter-socket hop on every H2D copy and shares that
link with the other ranks doing the same; ``bind_to_gpu_numa`` restricts the process (and every thread it creates afterwards: the OpenMP
pool of the host runtime, the prefetch threads) to the GPU's ideal CPU set as NVML reports it, so first-touch and ``cudaHostAlloc``
allocations land on the near socket.  Call it first thing in a rank, before the host runtime spins up its threads.

Reference: SessionGroup's ``cpusets`` / ``SESSION_GROUP_CPUSET`` (docs/docs_en/SessionGroup.md) pin serving sessions; the launcher
(``distribute/launch.py``) leaves training ranks unpinned."""
from __future__ import annotations

import os
from typing import List, Optional


def gpu_cpu_set(device_index: int) -> Optional[List[int]]:
    """CPUs NVML calls ideal for the GPU (``nvmlDeviceGetCpuAffinity``), or None when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        if os.environ.get("CUDA_VISIBLE_DEVICES"):                     # CUDA ordinals are remapped: resolve through the UUID
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
            except Exception:
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n = os.cpu_count() or 1
        words = (n + 63) // 64
        mask = list(pynvml.nvmlDeviceGetCpuAffinity(h, words))
        return [i for i in range(n) if (int(mask[i // 64]) >> (i % 64)) & 1]
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int, min_cpus: int = 8) -> Optional[List[int]]:
    """Restrict this process to (ideal CPUs of the GPU) & (CPUs it may already use).  No-op -- returns None -- when NVML is missing, the
    platform has no ``sched_setaffinity``, ``DEEPREC_NUMA_BIND=0``, or fewer than ``min_cpus`` CPUs would remain (containers with a narrow
    cpuset: better unpinned than starved)."""
    if os.environ.get("DEEPREC_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    ideal = gpu_cpu_set(device_index)
    if not ideal:
        return None
    try:
        allowed = os.sched_getaffinity(0)
        target = allowed & set(ideal)
        if len(target) < min_cpus or target == allowed:
            return None if target != allowed else sorted(target)
        os.sched_setaffinity(0, target)
        return sorted(target)
    except OSError:
        return None
