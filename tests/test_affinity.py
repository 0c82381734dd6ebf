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


def test_reference_env_flag_registry(monkeypatch):
    """config.REFERENCE_ENV_FLAGS: every 'honoured' switch is really read somewhere in the package; env_report lists what is set."""
    import subprocess
    from deeprec_b200 import config
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, (status, _note) in config.REFERENCE_ENV_FLAGS.items():
        assert status in ("honoured", "always", "n/a")
        if status == "honoured":
            hits = subprocess.run(["grep", "-rl", name, os.path.join(root, "deeprec_b200"), "--include=*.py", "--include=*.cc", "--include=*.h", "--include=*.cu"],
                                  capture_output=True, text=True).stdout.split()
            assert [h for h in hits if not h.endswith("config.py")], f"{name} is listed as honoured but nothing reads it"
    monkeypatch.setenv("ENABLE_MPS", "1"); monkeypatch.setenv("TF_GPU_VMEM", "1")
    rep = config.env_report(warn=False)
    assert rep["ENABLE_MPS"]["status"] == "n/a" and rep["TF_GPU_VMEM"]["status"] == "honoured"
