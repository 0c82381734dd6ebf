"""utils/affinity.py: NUMA binding is a guarded no-op without NVML / a GPU and never leaves the process with a starved CPU set."""
import os

from deeprec_b200.utils import affinity


def test_bind_is_a_noop_without_nvml_and_respects_the_switch(monkeypatch):
    before = os.sched_getaffinity(0)
    assert affinity.bind_to_gpu_numa(0) is None or os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
    monkeypatch.setenv("DEEPREC_NUMA_BIND", "0")
    monkeypatch.setattr(affinity, "gpu_cpu_set", lambda i: sorted(before)[:1])
    assert affinity.bind_to_gpu_numa(0) is None and os.sched_getaffinity(0) == before


def test_bind_intersects_with_the_allowed_set_and_keeps_a_minimum(monkeypatch):
    before = os.sched_getaffinity(0)
    cpus = sorted(before)
    try:
        monkeypatch.setattr(affinity, "gpu_cpu_set", lambda i: cpus[:1] + [10 ** 6])     # one usable CPU: fewer than min_cpus -> unpinned
        assert affinity.bind_to_gpu_numa(0) is None and os.sched_getaffinity(0) == before
        if len(cpus) >= 4:
            half = cpus[: len(cpus) // 2]
            monkeypatch.setattr(affinity, "gpu_cpu_set", lambda i: half + [10 ** 6])
            assert affinity.bind_to_gpu_numa(0, min_cpus=2) == half and os.sched_getaffinity(0) == set(half)
    finally:
        os.sched_setaffinity(0, before)
