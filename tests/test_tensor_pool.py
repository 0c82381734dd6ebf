"""TensorPool (planned allocator): collect -> plan -> serve, misses and re-planning, tensors backed by pool memory."""
import gc

import torch

from deeprec_b200.utils.memory import HostTensorPool


def _step(pool, shapes):
    ts = [pool.empty(s) for s in shapes]
    for i, t in enumerate(ts):
        t.fill_(float(i))
    ok = all(bool((t == float(i)).all()) for i, t in enumerate(ts))      # blocks do not overlap
    ptrs = sorted(t.data_ptr() for t in ts if t.numel() * 4 >= 4096)       # the pooled tensors only (tiny ones bypass the pool -> malloc)
    del ts, t
    gc.collect()
    pool.step_end()
    return ok, ptrs


def test_size_classes_bound_fragmentation():
    pool = HostTensorPool()
    for n in (256, 257, 320, 321, 4096, 4097, 5000, 1 << 20, (1 << 20) + 1, 123456789):
        cb = pool.class_bytes(n)
        assert n <= cb <= max(256, int(n * 1.26)), (n, cb)


def test_collect_plan_serve_and_replan():
    pool = HostTensorPool(small_threshold=4096, collect_steps=2, replan_misses=3)
    shapes = [(64, 128), (64, 128), (300, 40), (8,)]                       # two 32 KB, one 48 KB, one tiny (bypasses the pool)
    for _ in range(2):
        ok, _ = _step(pool, shapes)
        assert ok
    s = pool.stats()
    assert s["phase"] == 1 and s["pool_hits"] == 0 and s["small_bypass"] == 2
    assert s["pool_bytes"] == 2 * pool.class_bytes(64 * 128 * 4) + pool.class_bytes(300 * 40 * 4)
    base = s["backend_allocs"]
    seen = None
    for _ in range(5):
        ok, ptrs = _step(pool, shapes)
        assert ok
        seen = seen or ptrs
        assert ptrs == seen                                                  # the same pool blocks every step
    s = pool.stats()
    assert s["pool_hits"] == 15 and s["pool_misses"] == 0 and s["live_pool_blocks"] == 0
    assert s["backend_allocs"] == base + 5                                  # only the tiny bypass allocations reach malloc
    # the workload grows: one more 32 KB buffer per step -> misses -> a re-plan absorbs it
    bigger = shapes + [(64, 128)]
    for _ in range(3):
        assert _step(pool, bigger)[0]
    s2 = pool.stats()
    assert s2["pool_misses"] == 3 and s2["replans"] == 1 and s2["pool_bytes"] == s["pool_bytes"] + pool.class_bytes(64 * 128 * 4)
    assert _step(pool, bigger)[0]
    assert pool.stats()["pool_misses"] == 3                                # served from the grown pool now


def test_views_keep_the_block_alive():
    pool = HostTensorPool(small_threshold=256, collect_steps=1)
    pool.empty((1024,)); gc.collect(); pool.step_end()
    t = pool.empty((1024,))
    v = t[10:20]
    t.fill_(3.0)
    p = t.data_ptr()
    del t
    gc.collect()
    assert pool.stats()["live_pool_blocks"] == 1                           # the view still owns the block
    u = pool.empty((1024,))
    assert u.data_ptr() != p and bool((v == 3.0).all())
    del v, u
    gc.collect()
    assert pool.stats()["live_pool_blocks"] == 0


def test_statistic_step_env_switches(monkeypatch):
    """START_STATISTIC_STEP: allocations of the first steps (initialisation) do not enter the plan; STABLE/MAX_STATISTIC_STEP set the
    collection window; ENABLE_MEMORY_OPTIMIZATION=0 turns the pool into a pass-through."""
    from deeprec_b200.utils.memory import HostTensorPool
    monkeypatch.setenv("START_STATISTIC_STEP", "2"); monkeypatch.setenv("STABLE_STATISTIC_STEP", "2"); monkeypatch.setenv("MAX_STATISTIC_STEP", "50")
    pool = HostTensorPool.from_env()
    for step in range(6):
        if step < 2:
            big = [pool.empty((1 << 20,)) for _ in range(3)]            # 3 x 4 MiB only during initialisation
            del big
        a = pool.empty((100_000,)); b = pool.empty((100_000,))
        del a, b
        pool.step_end()
        st = pool.stats()
        assert st["phase"] == (1 if step >= 3 else 0), (step, st)      # window = steps 2, 3 -> planned at the end of step index 3
    st = pool.stats()
    assert st["pool_hits"] >= 4 and 2 * pool.class_bytes(400_000) <= st["pool_bytes"] < (4 << 20)     # the 4 MiB blocks were never planned
    monkeypatch.setenv("ENABLE_MEMORY_OPTIMIZATION", "0")
    off = HostTensorPool.from_env()
    for _ in range(5):
        x = off.empty((100_000,)); del x; off.step_end()
    assert off.stats()["pool_hits"] == 0 and off.stats()["pool_bytes"] == 0
