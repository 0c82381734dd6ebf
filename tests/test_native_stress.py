"""Builds tests/native/host_stress.cc against the host runtime sources with ThreadSanitizer (and AddressSanitizer) and runs it:
the race-detection tier for the lock-free host structures."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRCS = [os.path.join(ROOT, "deeprec_b200", "csrc", "host", f) for f in ("host_engine.cc", "io_runtime.cc", "ssd_store.cc")]


def _build_and_run(tmp_path, sanitizer):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / f"host_stress_{sanitizer}")
    cmd = ["g++", "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", "-pthread",
           os.path.join(ROOT, "tests", "native", "host_stress.cc"), *SRCS, "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip(f"-fsanitize={sanitizer} unsupported here")
    assert b.returncode == 0, b.stderr[-3000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 second_deadlock_stack=1", ASAN_OPTIONS="detect_leaks=0")
    run = [exe, str(tmp_path / "ssd")]
    if sanitizer == "thread" and shutil.which("setarch"):
        run = ["setarch", "-R"] + run            # TSAN + ASLR: glibc's tpp.c assertion / shadow-memory mapping failures are environment noise
    for attempt in range(3):
        r = subprocess.run(run, capture_output=True, text=True, timeout=900, env=env)
        if r.returncode == 0 or not ("tpp.c" in r.stderr or "unexpected memory mapping" in r.stderr):
            break
    else:
        pytest.skip("ThreadSanitizer runtime is not usable in this environment (glibc / ASLR incompatibility)")
    assert r.returncode == 0 and "HOST_STRESS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-6000:])


def test_host_runtime_under_thread_sanitizer(tmp_path):
    _build_and_run(tmp_path, "thread")


def test_host_runtime_under_address_sanitizer(tmp_path):
    _build_and_run(tmp_path, "address")


@pytest.mark.parametrize("sanitizer", ["address,undefined", "thread"])
def test_codec_fuzz_and_tensor_pool_under_sanitizers(tmp_path, sanitizer):
    """Protobuf request codec under random / truncated / bit-flipped input and the TensorPool under concurrent use."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "codec_pool_fuzz")
    cmd = ["g++", "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer", "-pthread",
           os.path.join(ROOT, "tests", "native", "codec_pool_fuzz.cc"), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip(f"-fsanitize={sanitizer} unsupported here")
    assert b.returncode == 0, b.stderr[-3000:]
    run = ["setarch", "-R", exe] if sanitizer == "thread" and shutil.which("setarch") else [exe]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1", ASAN_OPTIONS="detect_leaks=1")
    for attempt in range(3):
        r = subprocess.run(run, capture_output=True, text=True, timeout=900, env=env)
        if r.returncode == 0 or not ("tpp.c" in r.stderr or "unexpected memory mapping" in r.stderr):
            break
    else:
        pytest.skip("sanitizer runtime is not usable in this environment")
    assert r.returncode == 0 and "CODEC_POOL_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-6000:])


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_cpu_serving_runtime_under_sanitizers(tmp_path, sanitizer):
    """The CPU Processor served from 4 client threads while deltas and full versions are published underneath it."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "cpu_serving_stress")
    host = os.path.join(ROOT, "deeprec_b200", "csrc", "host")
    cmd = ["g++", "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", "-pthread",
           os.path.join(ROOT, "tests", "native", "cpu_serving_stress.cc"), os.path.join(host, "cpu_serving.cc"), os.path.join(host, "host_engine.cc"), os.path.join(host, "redis_store.cc"), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip(f"-fsanitize={sanitizer} unsupported here")
    assert b.returncode == 0, b.stderr[-3000:]
    run = ["setarch", "-R", exe] if sanitizer == "thread" and shutil.which("setarch") else [exe]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 second_deadlock_stack=1", ASAN_OPTIONS="detect_leaks=1")
    for attempt in range(3):
        work = str(tmp_path / f"models{attempt}")
        r = subprocess.run(run + [work], capture_output=True, text=True, timeout=900, env=env)
        if r.returncode == 0 or not ("tpp.c" in r.stderr or "unexpected memory mapping" in r.stderr):
            break
    else:
        pytest.skip("sanitizer runtime is not usable in this environment")
    assert r.returncode == 0 and "CPU_SERVING_STRESS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-6000:])


@pytest.mark.parametrize("sanitizer", ["thread", "address"])
def test_native_ps_data_plane_under_sanitizers(tmp_path, sanitizer):
    """csrc/host/ps_server.cc: four workers x (pull + push connections) against one server over real TCP while the scaling fence toggles; every
    accepted push applied exactly once, STALE answers keep the stream in sync, the fence drains, stop() joins with connections open."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / f"ps_stress_{sanitizer}")
    host = os.path.join(ROOT, "deeprec_b200", "csrc", "host")
    cmd = ["g++", "-O1", "-g", "-std=c++17", f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer", "-pthread", os.path.join(ROOT, "tests", "native", "ps_stress.cc"),
           os.path.join(host, "host_engine.cc"), os.path.join(host, "ps_server.cc"), "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if b.returncode != 0 and "sanitize" in b.stderr:
        pytest.skip(f"-fsanitize={sanitizer} unsupported here")
    assert b.returncode == 0, b.stderr[-3000:]
    run = ["setarch", "-R", exe] if sanitizer == "thread" and shutil.which("setarch") else [exe]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 second_deadlock_stack=1", ASAN_OPTIONS="detect_leaks=1")
    for attempt in range(3):
        r = subprocess.run(run, capture_output=True, text=True, timeout=900, env=env)
        if r.returncode == 0 or not ("tpp.c" in r.stderr or "unexpected memory mapping" in r.stderr):
            break
    else:
        pytest.skip("sanitizer runtime is not usable in this environment (glibc / ASLR incompatibility)")
    assert r.returncode == 0 and "PS_STRESS_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-6000:])
