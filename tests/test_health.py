"""Failure detection / fault injection / restart-and-resume."""
import os
import subprocess
import sys
import time

import torch

from deeprec_b200.utils import FaultInjector, HeartbeatMonitor, InjectedFault, StepWatchdog

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_watchdog_fires_once_per_stall_and_rearms():
    fired = []
    with StepWatchdog(0.3, on_stall=lambda idle: fired.append(idle), poll_s=0.05) as wd:
        for _ in range(5):
            time.sleep(0.05); wd.tick()
        assert not fired
        time.sleep(0.6)
        assert len(fired) == 1 and fired[0] > 0.3
        wd.tick(); time.sleep(0.6)
        assert len(fired) == 2


def test_heartbeat_detects_a_silent_rank(tmp_path):
    from torch.distributed import FileStore
    path = str(tmp_path / "store")
    a = HeartbeatMonitor(FileStore(path, 2), 0, 2, interval_s=0.05, timeout_s=0.4)
    b = HeartbeatMonitor(FileStore(path, 2), 1, 2, interval_s=0.05, timeout_s=0.4)
    time.sleep(0.3)
    assert a.dead_ranks() == [] and b.dead_ranks() == []
    b.close()                                    # rank 1 stops beating
    time.sleep(0.2); a.dead_ranks(); time.sleep(0.6)
    assert a.dead_ranks() == [1]
    a.close()


def test_fault_injector_spec():
    f = FaultInjector("step=3,kind=exception,rank=1", rank=0)
    f.maybe_fail(3)                              # other rank: nothing
    f = FaultInjector("step=3,kind=exception", rank=0)
    f.maybe_fail(2)
    try:
        f.maybe_fail(3)
        assert False
    except InjectedFault:
        pass
    f.maybe_fail(3)                              # fires once


def test_training_job_killed_mid_run_resumes_from_checkpoints(tmp_path):
    """The job dies at step 5 (hard exit, like a node failure); the restarted job restores the last full checkpoint + the
    incremental chain and carries on to the end (PS-failover semantics, docs Incremental-Checkpoint.md)."""
    cmd = [sys.executable, "-m", "deeprec_b200.models.train", "--model", "wdl", "--steps", "8", "--batch_size", "64", "--device", "cpu",
           "--checkpoint", str(tmp_path), "--save_steps", "2", "--log_every", "1"]
    env = dict(os.environ, DEEPREC_FAULT="step=5,kind=exit", PYTHONPATH=ROOT)
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r1.returncode == 13, r1.stderr[-2000:]
    assert "global_step 4" in r1.stdout and "global_step 6" not in r1.stdout
    env.pop("DEEPREC_FAULT")
    r2 = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "restored from" in r2.stdout and "global step 4" in r2.stdout          # last full checkpoint before the crash
    assert "global_step 5 " in r2.stdout and "global_step 12" in r2.stdout          # continued: 8 more steps from 4
