"""All-gather -> lookup -> reduce-scatter embedding as fused peer-memory kernels (csrc/cuda/ag_embedding.cu, parallel/ag_embedding.py) on the
CUDA-on-CPU emulation, ranks as threads: multi-hot sparse batches, sum / mean combiners, three training steps against a global fp32 oracle
(forward: combined rows of every sample; backward: one Adagrad update per distinct key with the gradient summed over every entry of every rank).

Reference dataflow: SOK v1 DistributedEmbedding (all_gather_dispatcher.cu:128-140, reduce_scatter_dispatcher.cu:42-84)."""
import math
import os
import threading

import pytest
import torch

from deeprec_b200 import _native

pytestmark = [pytest.mark.timeout(900)]
DEV = torch.device("cpu")
D = 16


def _table(owner):
    from deeprec_b200._native import EvConfig
    from deeprec_b200.ops.device_table import DeviceTable, get_context
    ctx = get_context(DEV, D, owner=owner)
    c = EvConfig()
    c.dim, c.num_slots, c.has_scalars = D, 1, 0
    c.init_capacity = 2048
    c.filter_type, c.filter_freq, c.bloom_counter_bits = 0, 0, 32
    c.steps_to_live, c.l2_weight_threshold = 0, -1.0
    c.default_value_dim, c.default_value_no_permission = 4096, 0.0
    c.record_freq = c.record_version = 1
    c.storage_type = 1
    c.slot_init[0] = 0.1
    dm = torch.empty(4096, D).normal_(0.0, 1.0 / math.sqrt(D), generator=torch.Generator().manual_seed(7))
    return ctx, DeviceTable(c, dm, DEV, capacity=1 << 12, row_capacity=1 << 11, owner=owner)


def _rank_main(rank, W, shared, steps, batches, douts, combiner, B, cap, out, errors):
    try:
        from deeprec_b200._native import OptHyper, ptr
        from deeprec_b200.optim.optimizers import OPT_ADAGRAD
        from deeprec_b200.parallel.ag_embedding import AllGatherEmbedding
        from deeprec_b200.parallel.emu_comm import EmuComm
        with _native.cuda_emulation():
            comm = EmuComm(shared, rank) if W > 1 else None
            ctx, table = _table(owner=8200 + rank)
            ag = AllGatherEmbedding(DEV, rank, W, table, B, cap, comm=comm)
            hp = OptHyper(); hp.kind, hp.lr, hp.init_accum = OPT_ADAGRAD, 0.1, 0.1
            ctx.set_hyper(hp)
            outs = []
            for s in range(steps):
                vals, rows = batches[s][rank]
                ag.load_ids(vals, rows)
                ag.lookup(ctx, train=True)
                outs.append(ag.reduce(combiner).clone())
                ag.stage_grad(douts[s][rank])
                ag.grad(ctx)
                rc = ag.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), D, ptr(ctx.hp_dev), W * cap, 1, None)
                assert rc == 0
                ag.step_end()
            out[rank] = (table, outs)
            if comm is not None:
                comm.host_barrier()
    except BaseException as e:                                  # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            shared.barrier.abort()
        except Exception:
            pass
        raise


@pytest.mark.parametrize("W,combiner", [(2, "mean")] if os.environ.get("DEEPREC_EMU_QUICK") == "1" else [(1, "sum"), (2, "mean"), (3, "sum")])
def test_all_gather_embedding_ranks_as_threads_match_the_global_oracle(W, combiner):
    from deeprec_b200.parallel.emu_comm import EmuWorld
    torch.manual_seed(20 + W)
    B, steps, cap, card = 48, 3, 256, 300
    batches, douts = [], []
    for s in range(steps):
        per_rank = []
        for r in range(W):
            n_per = torch.randint(1, 5, (B,))                       # 1..4 entries per sample (multi-hot)
            rows = torch.repeat_interleave(torch.arange(B), n_per)
            per_rank.append((torch.randint(0, card, (rows.numel(),)), rows))
        batches.append(per_rank)
        douts.append([torch.randn(B, D) for _ in range(W)])
    shared, out, errors = EmuWorld(W), {}, []
    with _native.cuda_emulation():
        pass
    threads = [threading.Thread(target=_rank_main, args=(r, W, shared, steps, batches, douts, combiner, B, cap, out, errors)) for r in range(W)]
    [t.start() for t in threads]
    [t.join(timeout=800) for t in threads]
    assert not errors and len(out) == W, errors
    dm = torch.empty(4096, D).normal_(0.0, 1.0 / math.sqrt(D), generator=torch.Generator().manual_seed(7))
    ref = {}
    for s in range(steps):
        gsum = {}
        for r in range(W):
            vals, rows = batches[s][r]
            cur = torch.stack([ref[int(k)][0] if int(k) in ref else dm[int(k) % 4096] for k in vals.tolist()])
            want = torch.zeros(B, D).index_add_(0, rows, cur)
            cnt = torch.bincount(rows, minlength=B).clamp(min=1).float()
            scale = (1.0 / cnt) if combiner == "mean" else torch.ones(B)
            want = want * scale.unsqueeze(1)
            got = out[r][1][s]
            assert (got - want).abs().max().item() < 1e-4, (s, r, float((got - want).abs().max()))
            g = (douts[s][r] * scale.unsqueeze(1))[rows]
            for k, gg in zip(vals.tolist(), g):
                gsum[k] = gsum.get(k, 0) + gg
        for k, gg in gsum.items():
            w, a = ref.get(k, (dm[k % 4096].clone(), torch.full((D,), 0.1)))
            a = a + gg * gg
            ref[k] = (w - 0.1 * gg / a.sqrt(), a)
    with _native.cuda_emulation():
        keys = torch.tensor(sorted(ref.keys()))
        want = torch.stack([ref[int(k)][0] for k in keys.tolist()])
        freq = torch.stack([out[r][0].get_freq(keys) for r in range(W)])
        assert ((freq > 0).sum(0) == 1).all()
        owner = (freq > 0).float().argmax(0)
        rows = torch.stack([out[r][0].lookup(keys) for r in range(W)])
        got = rows[owner, torch.arange(keys.numel())]
        assert (got - want).abs().max().item() < 1e-4
        total = sum(batches[s][r][0].numel() for s in range(steps) for r in range(W))
        assert int(freq.sum()) == total
