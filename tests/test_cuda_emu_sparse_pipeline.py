"""The unique-first model-parallel embedding pipeline (csrc/cuda/sparse_pipeline.cu + sp_sync.cuh + the device-table kernels) on the CUDA-on-CPU
emulation: world 1 against the torch oracles (the GPU test's assertions), and world 2 / 3 with the ranks as THREADS whose kernels talk through
each other's buffers exactly as the GPUs do over NVLink -- dedup -> bucket per owner -> owners probe / insert and push rows -> requesters gather
-> gradient pre-reduction -> owners pull and apply Adagrad -- against a global fp32 oracle.  Under DEEPREC_EMU_SANITIZE=thread the flag protocol
(release / acquire flags, last-block signalling, epoch counters) is race-checked by ThreadSanitizer (tests/test_cuda_emu_sanitizers.py)."""
import ctypes as C
import math
import os
import threading

import pytest
import torch

from deeprec_b200 import _native

pytestmark = [pytest.mark.timeout(900)]
DEV = torch.device("cpu")


def _mk_tables(dims, cards, owner):
    from deeprec_b200._native import EvConfig
    from deeprec_b200.ops.device_table import DeviceTable, get_context
    ctx = get_context(DEV, dims, owner=owner)
    tables = []
    for t, card in enumerate(cards):
        c = EvConfig()
        c.dim, c.num_slots, c.has_scalars = dims, 1, 0
        c.init_capacity = card
        c.filter_type, c.filter_freq = 0, 0
        c.bloom_counter_bits = 32
        c.steps_to_live, c.l2_weight_threshold = 0, -1.0
        c.default_value_dim, c.default_value_no_permission = 4096, 0.0
        c.record_freq = c.record_version = 1
        c.storage_type = 1
        c.slot_init[0] = 0.1
        g = torch.Generator().manual_seed(7 + t)
        dm = torch.empty(4096, dims).normal_(0.0, 1.0 / math.sqrt(dims), generator=g)
        tables.append(DeviceTable(c, dm, DEV, capacity=1 << 12, row_capacity=1 << 11, owner=owner))
    return ctx, tables


def test_dedup_inverse_and_counts_on_the_emulation():
    from deeprec_b200.parallel.sparse_pipeline import SparsePipeline
    torch.manual_seed(0)
    B, col_table = 300, [0, 1, 1, 2]                        # columns 1 and 2 share table 1 (sequence-style features)
    with _native.cuda_emulation():
        sp = SparsePipeline(DEV, 0, 1, col_table, 3, B, 16, pad_key=-1)
        ids = torch.stack([torch.randint(0, 40, (B,)), torch.randint(0, 300, (B,)), torch.randint(100, 400, (B,)), torch.randint(0, 5, (B,))])
        ids[1, ::7] = -1                                        # padding
        for rep in range(2):                                    # second round: the reset must have left the scratch clean
            sp.dedup(ids)
            scr = sp.scr_buf.tensor(torch.int64, (sp.Htot, 2))
            inv = sp.inv[:, : len(col_table)].t().long()        # [C][B]
            pad = ids == -1
            assert (inv[pad] == -1).all() and (inv[~pad] >= 0).all()
            assert torch.equal(scr[inv[~pad], 0], ids[~pad]), "inv must point at the slot holding the key"
            n_unique = 0
            for t in range(3):
                cols = [c for c, tt in enumerate(col_table) if tt == t]
                k = torch.cat([ids[c][ids[c] != -1] for c in cols])
                u, cnt = torch.unique(k, return_counts=True)
                n_unique += u.numel()
                gs = torch.cat([inv[c][ids[c] != -1] for c in cols])
                assert torch.unique(gs).numel() == u.numel()
                got = {int(a): int(b) & 0xFFFFFFFF for a, b in zip(scr[torch.unique(gs), 0].tolist(), scr[torch.unique(gs), 1].tolist())}
                assert got == {int(a): int(b) for a, b in zip(u.tolist(), cnt.tolist())}
                assert int(sp.bcnt[t, 0]) == u.numel()
            assert sp.unique_count() == n_unique
            sp.reset(); sp.step_end()
            assert int(sp.bcnt.sum()) == 0
            assert int((scr[:, 0] != -(1 << 63)).sum()) == 0, "reset must free every touched scratch slot"


def _rank_main(rank, W, shared, steps, ids_all, grads_all, D, cards, out, errors):
    """One emulated rank: its own tables (the keys it owns), its own batch, the shared-address-space peers."""
    try:
        from deeprec_b200._native import OptHyper, ptr
        from deeprec_b200.optim.optimizers import OPT_ADAGRAD
        from deeprec_b200.parallel.emu_comm import EmuComm
        from deeprec_b200.parallel.sparse_pipeline import SparsePipeline
        with _native.cuda_emulation():
            comm = EmuComm(shared, rank) if W > 1 else None
            ctx, tables = _mk_tables(D, cards, owner=7000 + rank)
            tmap = torch.tensor([t.gid for t in tables], dtype=torch.int32)
            B = ids_all[0][rank].shape[1]
            sp = SparsePipeline(DEV, rank, W, list(range(len(cards))), len(cards), B, D, comm=comm)
            ctx.ensure(len(cards) * B * W)
            hp = OptHyper(); hp.kind, hp.lr, hp.init_accum = OPT_ADAGRAD, 0.1, 0.1
            ctx.set_hyper(hp)
            gathered = []
            for step in range(steps):
                ids, g = ids_all[step][rank], grads_all[step][rank]
                sp.dedup(ids)                                   # raises DEDUP on every peer
                sp.lookup(ctx, tmap, True)                      # waits DEDUP of every source, pushes rows, raises ROWS
                rows = torch.empty(B, len(cards), D, dtype=torch.bfloat16)
                sp.gather(rows)                                 # waits ROWS of every owner
                gathered.append(rows.float().clone())
                sp.segsum(g)                                    # pre-reduce per distinct key, raises GRAD
                sp.reset()
                sp.grad(ctx, tmap)                              # waits GRAD of every source, pulls their rows
                rc = sp.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), D, ptr(ctx.hp_dev),
                                                 len(cards) * B * W, 1, None)
                assert rc == 0
                sp.step_end()
                if comm is not None:
                    comm.host_barrier()
            out[rank] = (tables, gathered)
            if comm is not None:
                comm.host_barrier()
    except BaseException as e:                                  # noqa: BLE001 -- surfaced by the main thread
        errors.append((rank, repr(e)))
        try:
            shared.barrier.abort()
        except Exception:
            pass
        raise


@pytest.mark.parametrize("W", [2] if os.environ.get("DEEPREC_EMU_QUICK") == "1" else [1, 2, 3])
def test_pipeline_ranks_as_threads_match_the_global_oracle(W):
    from deeprec_b200.parallel.emu_comm import EmuWorld
    torch.manual_seed(10 + W)
    B, D, cards, steps = 96, 16, [37, 900], 3
    ids_all = [[torch.stack([torch.randint(0, c, (B,)) for c in cards]) for _ in range(W)] for _ in range(steps)]
    grads_all = [[torch.randn(len(cards), B, D).bfloat16() for _ in range(W)] for _ in range(steps)]
    shared, out, errors = EmuWorld(W), {}, []
    with _native.cuda_emulation():          # build + load the library on the main thread (the rank threads must not fork a compiler)
        pass
    threads = [threading.Thread(target=_rank_main, args=(r, W, shared, steps, ids_all, grads_all, D, cards, out, errors)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=800)
    assert not errors and len(out) == W, errors
    # global fp32 oracle: per step, sum the gradients of every occurrence of a key on every rank, one Adagrad update per distinct key
    dms = [torch.empty(4096, D).normal_(0.0, 1.0 / math.sqrt(D), generator=torch.Generator().manual_seed(7 + t)) for t in range(len(cards))]
    ref = [dict() for _ in cards]
    for step in range(steps):
        for t in range(len(cards)):
            # rows every requester gathered at this step == the parameters BEFORE this step's update
            for r in range(W):
                keys = ids_all[step][r][t]
                want = torch.stack([ref[t][int(k)][0] if int(k) in ref[t] else dms[t][int(k) % 4096] for k in keys.tolist()])
                got = out[r][1][step][:, t, :]
                assert (got - want).abs().max().item() < 1e-2, (step, t, r)
            keys = torch.cat([ids_all[step][r][t] for r in range(W)])
            g = torch.cat([grads_all[step][r][t].float() for r in range(W)])
            u, invu = torch.unique(keys, return_inverse=True)
            gsum = torch.zeros(u.numel(), D).index_add_(0, invu, g)
            for k, gg in zip(u.tolist(), gsum):
                w, a = ref[t].get(k, (dms[t][k % 4096].clone(), torch.full((D,), 0.1)))
                a = a + gg * gg
                ref[t][k] = (w - 0.1 * gg / a.sqrt(), a)
    with _native.cuda_emulation():
        for t in range(len(cards)):
            keys = torch.tensor(sorted(ref[t].keys()))
            want = torch.stack([ref[t][int(k)][0] for k in keys.tolist()])
            freq = torch.stack([out[r][0][t].get_freq(keys) for r in range(W)])           # [W, n]: exactly one owner holds each key
            assert ((freq > 0).sum(0) == 1).all(), "every key lives on exactly one owner"
            owner = (freq > 0).float().argmax(0)
            rows = torch.stack([out[r][0][t].lookup(keys) for r in range(W)])              # [W, n, D]
            got = rows[owner, torch.arange(keys.numel())]
            assert (got - want).abs().max().item() < 1e-4, t
            total = torch.cat([ids_all[s][r][t] for s in range(steps) for r in range(W)])
            assert int(freq.sum()) == total.numel(), "frequency counts every occurrence exactly once across the owners"
