"""csrc/cuda/ag_embedding.cu on hardware at world 1 (the multi-rank protocol is covered by the emulation test tests/test_cuda_emu_ag_embedding.py
with ranks as threads): multi-hot lookup + mean combine + Adagrad against the fp32 oracle.  Written without GPU access in the last session of
round 2; sorts last."""
import math

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_all_gather_embedding_single_gpu_matches_oracle():
    from deeprec_b200._native import EvConfig, OptHyper, ptr, stream_ptr
    from deeprec_b200.ops.device_table import DeviceTable, get_context
    from deeprec_b200.optim.optimizers import OPT_ADAGRAD
    from deeprec_b200.parallel.ag_embedding import AllGatherEmbedding
    dev = torch.device("cuda", 0)
    D, B, cap, card = 16, 512, 4096, 2000
    ctx = get_context(dev, D, owner=8801)
    c = EvConfig()
    c.dim, c.num_slots, c.has_scalars, c.init_capacity = D, 1, 0, 8192
    c.bloom_counter_bits, c.l2_weight_threshold, c.default_value_dim = 32, -1.0, 4096
    c.record_freq = c.record_version = 1
    c.storage_type = 1
    c.slot_init[0] = 0.1
    dm = torch.empty(4096, D).normal_(0.0, 1.0 / math.sqrt(D), generator=torch.Generator().manual_seed(7))
    table = DeviceTable(c, dm, dev, capacity=1 << 14, row_capacity=1 << 13, owner=8801)
    ag = AllGatherEmbedding(dev, 0, 1, table, B, cap)
    hp = OptHyper(); hp.kind, hp.lr, hp.init_accum = OPT_ADAGRAD, 0.1, 0.1
    ctx.set_hyper(hp)
    torch.manual_seed(0)
    ref = {}
    for step in range(3):
        n_per = torch.randint(1, 5, (B,))
        rows = torch.repeat_interleave(torch.arange(B), n_per)
        vals = torch.randint(0, card, (rows.numel(),))
        dout = torch.randn(B, D)
        ag.load_ids(vals.to(dev), rows.to(dev))
        ag.lookup(ctx, train=True)
        out = ag.reduce("mean")
        cur = torch.stack([ref[int(k)][0] if int(k) in ref else dm[int(k) % 4096] for k in vals.tolist()])
        cnt = torch.bincount(rows, minlength=B).clamp(min=1).float()
        want = torch.zeros(B, D).index_add_(0, rows, cur) / cnt.unsqueeze(1)
        assert (out.cpu() - want).abs().max().item() < 1e-4
        ag.stage_grad(dout.to(dev))
        ag.grad(ctx)
        assert ag.lib.dr_cuda_sparse_apply(ptr(ctx.structs()), ptr(ctx.ulist), ptr(ctx.nuniq), ctx.ulist.numel(), ptr(ctx.gsum), D, ptr(ctx.hp_dev), cap, 1, stream_ptr()) == 0
        ag.step_end()
        torch.cuda.synchronize()
        g = (dout / cnt.unsqueeze(1))[rows]
        gsum = {}
        for k, gg in zip(vals.tolist(), g):
            gsum[k] = gsum.get(k, 0) + gg
        for k, gg in gsum.items():
            w, a = ref.get(k, (dm[k % 4096].clone(), torch.full((D,), 0.1)))
            a = a + gg * gg
            ref[k] = (w - 0.1 * gg / a.sqrt(), a)
    keys = torch.tensor(sorted(ref.keys()), device=dev)
    want = torch.stack([ref[int(k)][0] for k in keys.tolist()])
    assert (table.lookup(keys).cpu() - want).abs().max().item() < 1e-4
