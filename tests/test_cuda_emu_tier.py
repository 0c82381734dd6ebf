"""Native multi-tier storage (csrc/cuda/tier_kernels.cu + ops/tier_manager.py) on the CUDA-on-CPU emulation: a small "HBM" cache over the host
DRAM tier, driven exactly as FusedRecEngine drives it (prefetch the NEXT batch's ids, commit at the step boundary, then the step's sparse
pipeline: dedup -> probe / insert -> gradient -> Adagrad), must end with the same parameters as one big table -- while the manager's background
thread stages promotions and commits demotions concurrently with the emulated kernels.  Under DEEPREC_EMU_SANITIZE=thread this is the race check
of the mapped-pinned-memory / event / condition-variable protocol between the kernels and the EvictionManager thread."""
import ctypes as C
import math
import os

import pytest
import torch

from deeprec_b200 import _native

pytestmark = [pytest.mark.timeout(900)]
DEV = torch.device("cpu")


def _table(D, rows, cap, owner):
    from deeprec_b200._native import EvConfig
    from deeprec_b200.ops.device_table import DeviceTable, get_context
    ctx = get_context(DEV, D, owner=owner)
    c = EvConfig()
    c.dim, c.num_slots, c.has_scalars = D, 1, 0
    c.init_capacity = rows
    c.filter_type, c.filter_freq, c.bloom_counter_bits = 0, 0, 32
    c.steps_to_live, c.l2_weight_threshold = 0, -1.0
    c.default_value_dim, c.default_value_no_permission = 4096, 0.0
    c.record_freq = c.record_version = 1
    c.storage_type = 1
    c.slot_init[0] = 0.1
    dm = torch.empty(4096, D).normal_(0.0, 1.0 / math.sqrt(D), generator=torch.Generator().manual_seed(7))
    return ctx, DeviceTable(c, dm, DEV, capacity=cap, row_capacity=rows, owner=owner)


@pytest.mark.parametrize("strategy", [0] if os.environ.get("DEEPREC_EMU_QUICK") == "1" else [0, 1])
def test_small_cache_over_the_host_tier_trains_like_one_big_table(strategy):
    from deeprec_b200._native import OptHyper, ptr
    from deeprec_b200.ops.tier_manager import DeviceTierManager
    from deeprec_b200.optim.optimizers import OPT_ADAGRAD
    from deeprec_b200.parallel.sparse_pipeline import SparsePipeline
    B, D, steps, cache = 128, 16, 12, 256
    torch.manual_seed(3 + strategy)
    ids_all, grads_all = [], []
    for s in range(steps):
        lo = (s % 4) * 400                                           # rotating working set: 1600 distinct ids >> 256 cache rows
        ids_all.append(torch.randint(lo, lo + 400, (1, B)))
        grads_all.append(torch.randn(1, B, D).bfloat16())
    with _native.cuda_emulation():
        results = []
        for tiered in (False, True):
            rows = 4096 if not tiered else cache + 2 * B + 1024
            ctx, table = _table(D, rows, 1 << 13 if not tiered else 1 << 11, owner=9100 + 10 * strategy + int(tiered))
            mgr = DeviceTierManager(table, cache, strategy=strategy, max_batch_keys=1 << 12, evict_chunk=128) if tiered else None
            tmap = torch.tensor([table.gid], dtype=torch.int32)
            sp = SparsePipeline(DEV, 0, 1, [0], 1, B, D)
            ctx.ensure(B)
            hp = OptHyper(); hp.kind, hp.lr, hp.init_accum = OPT_ADAGRAD, 0.1, 0.1
            ctx.set_hyper(hp)
            if mgr:
                mgr.prefetch(ids_all[0].reshape(-1).contiguous())
            for s in range(steps):
                if mgr:
                    mgr.commit(s)                                        # promoted rows resident, cold rows demoted before the step's kernels
                ctx.set_step(s)
                sp.dedup(ids_all[s]); sp.lookup(ctx, tmap, True)
                sp.segsum(grads_all[s]); sp.reset(); sp.grad(ctx, tmap)
                assert sp.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), D, ptr(ctx.hp_dev), B, 1, None) == 0
                sp.step_end()
                if mgr and s + 1 < steps:
                    mgr.prefetch(ids_all[s + 1].reshape(-1).contiguous())     # one batch ahead
            probe = torch.arange(0, 1600)
            results.append((mgr.lookup(probe) if mgr else table.lookup(probe)).clone())
            if mgr:
                st = mgr.stats()
                assert st["demoted_rows"] > 0 and st["promoted_rows"] > 0 and st["evict_passes"] > 0, st
                assert st["hbm_rows"] <= cache + 2 * B + 1024 and table.overflowed() == 0
                mgr.close()
        assert torch.allclose(results[0], results[1], atol=1e-5), float((results[0] - results[1]).abs().max())


# ---- world > 1: every rank holds both tiers of the keys it owns; the prefetch is a fused id all-gather + probe over peer memory ----------------------
def _tier_rank_main(rank, W, shared, steps, ids_all, grads_all, D, cache, tiered, strategy, out, errors):
    try:
        from deeprec_b200._native import OptHyper, ptr
        from deeprec_b200.ops.tier_manager import DeviceTierManager
        from deeprec_b200.optim.optimizers import OPT_ADAGRAD
        from deeprec_b200.parallel.emu_comm import EmuComm
        from deeprec_b200.parallel.sparse_pipeline import SparsePipeline
        with _native.cuda_emulation():
            comm = EmuComm(shared, rank)
            B = ids_all[0][rank].shape[1]
            rows = 4096 if not tiered else cache + 2 * B * W + 1024
            ctx, table = _table(D, rows, 1 << 13 if not tiered else 1 << 11, owner=9300 + 100 * int(tiered) + 10 * strategy + rank)
            mgr = DeviceTierManager(table, cache, strategy=strategy, max_batch_keys=1 << 12, evict_chunk=128, comm=comm, ids_per_prefetch=B) if tiered else None
            tmap = torch.tensor([table.gid], dtype=torch.int32)
            sp = SparsePipeline(DEV, rank, W, [0], 1, B, D, comm=comm)
            ctx.ensure(B * W)
            hp = OptHyper(); hp.kind, hp.lr, hp.init_accum = OPT_ADAGRAD, 0.1, 0.1
            ctx.set_hyper(hp)
            if mgr:
                mgr.prefetch(ids_all[0][rank].reshape(-1).contiguous())
            for s in range(steps):
                if mgr:
                    mgr.commit(s)
                ctx.set_step(s)
                sp.dedup(ids_all[s][rank]); sp.lookup(ctx, tmap, True)
                got = torch.empty(B, 1, D, dtype=torch.bfloat16); sp.gather(got)
                sp.segsum(grads_all[s][rank]); sp.reset(); sp.grad(ctx, tmap)
                assert sp.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), D, ptr(ctx.hp_dev), B * W, 1, None) == 0
                comm.host_barrier()                 # stands in for the engines' dense all-reduce rendezvous (every owner has pulled my gradients before my next dedup zeroes them)
                sp.step_end()
                if mgr and s + 1 < steps:
                    mgr.prefetch(ids_all[s + 1][rank].reshape(-1).contiguous())        # MY next batch; the kernels read every peer's in place (no host barrier around it)
            comm.host_barrier()
            probe = torch.arange(0, 1600)
            res = (mgr.lookup(probe) if mgr else table.lookup(probe)).clone()
            st = None
            if mgr:
                st = mgr.stats()
                st["overflowed"] = table.overflowed()
                mgr.close()
            out[rank] = (res, st)
            comm.host_barrier()
    except BaseException as e:                                  # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            shared.barrier.abort()
        except Exception:
            pass
        raise


@pytest.mark.parametrize("W", [2] if os.environ.get("DEEPREC_EMU_QUICK") == "1" else [2, 3])
def test_owner_side_promotion_with_ranks_as_threads(W):
    """Each rank: 256-row cache over its own DRAM tier for the keys it owns; ids of a key arrive in EVERY rank's batches.  Must train exactly like
    W ranks with big single-tier tables."""
    import threading
    from deeprec_b200.checkpoint.engine_ckpt import sp_owner
    from deeprec_b200.parallel.emu_comm import EmuWorld
    B, D, steps, cache, strategy = 96, 16, 10, 256, W % 2
    torch.manual_seed(40 + W)
    ids_all, grads_all = [], []
    for s in range(steps):
        lo = (s % 4) * 400
        ids_all.append([torch.randint(lo, lo + 400, (1, B)) for _ in range(W)])
        grads_all.append([torch.randn(1, B, D).bfloat16() for _ in range(W)])
    with _native.cuda_emulation():
        pass
    results = []
    for tiered in (False, True):
        shared, out, errors = EmuWorld(W), {}, []
        threads = [threading.Thread(target=_tier_rank_main, args=(r, W, shared, steps, ids_all, grads_all, D, cache, tiered, strategy, out, errors)) for r in range(W)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=800)
        assert not errors and len(out) == W, errors
        results.append(out)
    probe = torch.arange(0, 1600)
    owner = sp_owner(probe, W)
    rows = [torch.stack([results[i][r][0] for r in range(W)])[owner, torch.arange(1600)] for i in range(2)]      # every key read from its owner
    assert torch.allclose(rows[0], rows[1], atol=1e-5), float((rows[0] - rows[1]).abs().max())
    for r in range(W):
        st = results[1][r][1]
        assert st["demoted_rows"] > 0 and st["promoted_rows"] > 0 and st["evict_passes"] > 0 and st["overflowed"] == 0, (r, st)
        assert st["dram_rows"] > 0 and st["hbm_rows"] <= cache + 2 * B * W + 1024
