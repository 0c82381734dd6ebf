"""GPU: device EmbeddingVariable vs the host engine oracle; fused group lookup; DLRMEngine vs fp32 PyTorch DLRM."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(name, dim, device, **kw):
    import deeprec_b200 as dr
    opt = dr.EmbeddingVariableOption(**kw)
    return dr.get_embedding_variable(name, dim, ev_option=opt, device=device, seed=7)


@pytest.mark.parametrize("opt_name", ["adagrad", "adam", "adamasync", "adamw", "ftrl", "gradientdescent", "adagraddecay"])
def test_device_table_matches_host(opt_name):
    import deeprec_b200 as dr
    from deeprec_b200.optim import make_optimizer, GlobalStep
    torch.manual_seed(0)
    evs = {d: _mk(f"tbl_{opt_name}_{d}", 16, d) for d in ("cpu", "cuda")}
    kw = dict(lr=0.05)
    if opt_name == "ftrl":
        kw.update(l1_regularization_strength=0.01, l2_regularization_strength=0.01)
    if opt_name == "adagraddecay":
        kw.update(accumulator_decay_step=2, accumulator_decay_rate=0.5)
    opts = {d: make_optimizer(opt_name, [], [evs[d]], global_step=GlobalStep(), **kw) for d in evs}
    for step in range(4):
        ids = torch.randint(0, 50, (200,))
        tgt = torch.randn(200, 16)
        for d in evs:
            e = evs[d].lookup(ids.to(d))
            loss = ((e - tgt.to(d)) ** 2).sum()
            loss.backward()
            opts[d].step()
    probe = torch.arange(0, 60)
    a = evs["cpu"].table.lookup(probe)
    b = evs["cuda"].table.lookup(probe.cuda()).cpu()
    assert torch.allclose(a, b, atol=2e-4, rtol=2e-4), (a - b).abs().max()
    assert evs["cpu"].total_count() == evs["cuda"].total_count()
    assert torch.equal(evs["cpu"].get_frequency(probe), evs["cuda"].get_frequency(probe))


def test_device_counter_filter_and_eviction_and_snapshot():
    import deeprec_b200 as dr
    ev = _mk("flt", 8, "cuda", filter_option=dr.CounterFilter(3), evict_option=dr.GlobalStepEvict(2))
    opt = dr.optim.AdagradOptimizer([], [ev], lr=0.1, global_step=dr.optim.GlobalStep())
    ids = torch.tensor([1, 1, 2, 3, 3, 3], device="cuda")
    e = ev.lookup(ids)
    assert e.abs().max().item() == 0.0                       # un-admitted keys read default_value_no_permission
    e.sum().backward(); opt.step()
    assert ev.total_count() == 1                             # only key 3 reached freq 3
    assert ev.get_frequency(torch.tensor([1, 2, 3, 4])).tolist() == [2, 1, 3, 0]
    snap = ev.table.snapshot()
    assert snap["keys"].tolist() == [3] and sorted(snap["keys_filtered"].tolist()) == [1, 2]
    assert int(snap["partition_offset"][-1]) == 1
    # re-import into a fresh table with a partition filter
    ev2 = _mk("flt2", 8, "cuda", filter_option=dr.CounterFilter(3))
    ev2._set_slots(["accumulator"], [0.1], False)
    n = ev2.table.import_(snap["keys"], snap["rows"], snap["freqs"], snap["versions"])
    assert n == 1 and torch.allclose(ev2.table.lookup(torch.tensor([3])).cpu(), ev.table.lookup(torch.tensor([3])).cpu())
    # eviction: steps_to_live=2 -> untouched keys go at step > version + 2
    for _ in range(4):
        ev.lookup(torch.tensor([3, 3, 3], device="cuda")).sum().backward(); opt.step()
    removed = ev.table.shrink(int(opt.global_step))
    assert removed == 2 and ev.table.total_keys() == 1
    assert ev.get_frequency(torch.tensor([3])).item() >= 3


def test_device_table_growth():
    import deeprec_b200 as dr
    ev = _mk("grow", 8, "cuda", init_capacity=1024)
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=1.0, global_step=dr.optim.GlobalStep())
    total = 0
    for i in range(6):
        ids = torch.arange(i * 5000, (i + 1) * 5000, device="cuda")
        ev.lookup(ids).sum().backward(); opt.step()
        total += 5000
    assert ev.total_count() == total and ev.table.overflowed() == 0
    probe = torch.tensor([0, 4999, 29999], device="cuda")
    rows = ev.table.lookup(probe)
    dm = ev.default_matrix.cuda()
    assert torch.allclose(rows, dm[probe % dm.shape[0]] - 1.0, atol=1e-5)


def test_group_lookup_matches_python_path():
    import deeprec_b200 as dr
    torch.manual_seed(1)
    B = 64
    dev_evs = [_mk(f"g{i}", 16, "cuda") for i in range(3)]
    cpu_evs = [_mk(f"g{i}", 16, "cpu") for i in range(3)]
    gs = {d: dr.optim.GlobalStep() for d in ("cpu", "cuda")}
    o_dev = dr.optim.AdagradOptimizer([], dev_evs, lr=0.1, global_step=gs["cuda"])
    o_cpu = dr.optim.AdagradOptimizer([], cpu_evs, lr=0.1, global_step=gs["cpu"])
    for step in range(3):
        sps, ws = [], []
        for i in range(3):
            lens = torch.randint(0, 4, (B,))
            vals = torch.randint(0, 30, (int(lens.sum()),))
            off = torch.zeros(B + 1, dtype=torch.int64); off[1:] = torch.cumsum(lens, 0)
            w = torch.rand(vals.numel()) + 0.5 if i == 1 else None
            sps.append(dr.SparseIds.from_offsets(vals, off, w))
        combs = ["sum", "mean", "sqrtn"]
        out_dev = dr.group_embedding_lookup_sparse(dev_evs, [s.to("cuda") for s in sps], combs)
        out_cpu = [dr.embedding_lookup_sparse(e, s, None, c) for e, s, c in zip(cpu_evs, sps, combs)]
        for a, b in zip(out_dev, out_cpu):
            assert torch.allclose(a.cpu(), b, atol=1e-4), (a.cpu() - b).abs().max()
        tgt = [torch.randn(B, 16) for _ in range(3)]
        sum(((a - t.cuda()) ** 2).sum() for a, t in zip(out_dev, tgt)).backward(); o_dev.step()
        sum(((b - t) ** 2).sum() for b, t in zip(out_cpu, tgt)).backward(); o_cpu.step()
    probe = torch.arange(0, 30)
    for d, c in zip(dev_evs, cpu_evs):
        assert torch.allclose(d.table.lookup(probe).cpu(), c.table.lookup(probe), atol=1e-3)


def _oracle_from_engine(eng):
    """fp32 PyTorch DLRM carrying the engine's initial parameters."""
    from deeprec_b200.models.dlrm import DLRM
    cfg = eng.cfg
    m = DLRM(cfg.num_dense, cfg.cardinalities, cfg.embedding_dim, cfg.mlp_bot, cfg.mlp_top, use_ev=False, device="cuda",
             bn_eps=cfg.bn_eps, bn_momentum=cfg.bn_momentum)
    with torch.no_grad():
        lins = [l for l in m.bot if isinstance(l, torch.nn.Linear)]
        for L, lin in zip(eng.bot, lins):
            lin.weight.copy_(eng.p(L.name + "/kernel").view(L.N, L.Kp)[:, : L.K]); lin.bias.copy_(eng.p(L.name + "/bias"))
        lins = [l for l in m.top if isinstance(l, torch.nn.Linear)]
        for L, lin in zip(eng.top, lins):
            lin.weight.copy_(eng.p(L.name + "/kernel").view(L.N, L.Kp)[:, : L.K]); lin.bias.copy_(eng.p(L.name + "/bias"))
        m.logits.weight.copy_(eng.p("logits/kernel").view(1, -1)); m.logits.bias.copy_(eng.p("logits/bias")[:1])
        for t, emb in enumerate(m.tables):
            dm = eng.tables[t].default_matrix
            idx = torch.arange(emb.num_embeddings, device="cuda") % dm.shape[0]
            emb.weight.copy_(dm[idx])
    return m


@pytest.mark.parametrize("graph", [False, True])
def test_dlrm_engine_matches_fp32_oracle(graph):
    from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
    torch.manual_seed(0)
    cards = [50, 1000, 7, 300] + [97] * 22
    cfg = DLRMConfig(batch_size=1024, cardinalities=cards, optimizer="adagrad", learning_rate=0.05)
    eng = DLRMEngine(cfg)
    m = _oracle_from_engine(eng)
    dense_params = [p for n, p in m.named_parameters() if not n.startswith("tables")]
    o_dense = torch.optim.Adagrad(dense_params, lr=0.05, initial_accumulator_value=0.1, eps=0.0)
    o_emb = torch.optim.Adagrad(m.tables.parameters(), lr=0.05, initial_accumulator_value=0.1, eps=0.0)
    batches = []
    for s in range(4):
        dense = torch.rand(cfg.batch_size, 13, device="cuda") * 3
        ids = torch.stack([torch.randint(0, c, (cfg.batch_size,), device="cuda") for c in cards])
        labels = (torch.rand(cfg.batch_size, device="cuda") < 0.3).float()
        batches.append((dense, ids, labels))
    losses_e, losses_r = [], []
    eng.load_batch(*batches[0])
    if graph:
        eng.capture()          # runs one eager step on batch 0, then captures
        losses_e.append(eng.loss_value())
    for s, (dense, ids, labels) in enumerate(batches):
        if not (graph and s == 0):
            eng.load_batch(dense, ids, labels)
            eng.train_step()
            losses_e.append(eng.loss_value())
        loss = m.loss(dense, ids, labels)
        o_dense.zero_grad(); o_emb.zero_grad()
        loss.backward()
        o_dense.step(); o_emb.step()
        losses_r.append(loss.item())
    for a, b in zip(losses_e, losses_r):
        assert abs(a - b) < 0.02 * max(1.0, abs(b)), (losses_e, losses_r)
    assert losses_e[-1] < losses_e[0] + 0.05
    # updated parameters stay close to the fp32 oracle (bf16 activations => loose tolerance)
    lin0 = [l for l in m.bot if isinstance(l, torch.nn.Linear)][0]
    w_e = eng.p("mlp_bot_0/kernel").view(512, 16)[:, :13]
    assert (w_e - lin0.weight).abs().max().item() < 0.05
    # embedding rows of touched keys moved identically (same dedup semantics)
    t = 1
    keys = batches[-1][1][t][:64]
    rows_e = eng.tables[t].lookup(keys)
    rows_r = m.tables[t].weight[keys]
    assert (rows_e - rows_r).abs().max().item() < 0.05
    assert eng.tables[0].overflowed() == 0


def test_multi_tier_hbm_dram_matches_single_tier():
    """HBM cache (tiny) over DRAM: training results equal the all-HBM table; cold rows are demoted and promoted back."""
    import deeprec_b200 as dr
    torch.manual_seed(0)
    small = dr.StorageOption(dr.StorageType.HBM_DRAM, storage_size=(2048 * 2 * 16 * 4,), cache_strategy=dr.CacheStrategy.LFU)
    ev_mt = _mk("mt", 16, "cuda", storage_option=small)
    ev_ref = _mk("mt", 16, "cuda")
    gs1, gs2 = dr.optim.GlobalStep(), dr.optim.GlobalStep()
    o_mt = dr.optim.AdagradOptimizer([], [ev_mt], lr=0.1, global_step=gs1)
    o_ref = dr.optim.AdagradOptimizer([], [ev_ref], lr=0.1, global_step=gs2)
    from deeprec_b200.ops.multi_tier import MultiTierTable
    assert isinstance(ev_mt.table, MultiTierTable) and ev_mt.table.cache_rows == 2048
    for step in range(12):
        lo = (step % 4) * 1500                                  # working set rotates: 6000 distinct ids > 2048-row cache
        ids = torch.randint(lo, lo + 1500, (1200,), device="cuda")
        if step % 2 == 0:
            ev_mt.table.prefetch(ids)                              # staged-pipeline style prefetch + pin
        tgt = torch.randn(1200, 16, device="cuda")
        for ev, opt in ((ev_mt, o_mt), (ev_ref, o_ref)):
            ((ev.lookup(ids) - tgt) ** 2).sum().backward(); opt.step()
    probe = torch.arange(0, 6000, 7, device="cuda")
    a = ev_mt.table.lookup(probe).cpu(); b = ev_ref.table.lookup(probe).cpu()
    assert torch.allclose(a, b, atol=1e-4), (a - b).abs().max()
    st = ev_mt.table.cache_stats()
    assert st["misses"] > 0 and st["hbm_rows"] <= 2048 + 1024
    tiers = ev_mt.lookup_tier(torch.tensor([0, 5999, 10 ** 9]))
    assert tiers[2].item() == -1 and set(tiers[:2].tolist()) <= {0, 1}
    assert ev_mt.total_count() == ev_ref.total_count()
    snap = ev_mt.table.snapshot()
    assert snap["keys"].numel() == ev_ref.total_count()


def test_three_tier_hbm_dram_ssd(tmp_path):
    """HBM_DRAM_SSDHASH: tiny HBM cache over a tiny DRAM tier over the log-structured SSD store; results equal all-HBM training."""
    import deeprec_b200 as dr
    torch.manual_seed(0)
    rb = 2 * 16 * 4
    so = dr.StorageOption(dr.StorageType.HBM_DRAM_SSDHASH, storage_path=str(tmp_path), storage_size=(2048 * rb, 1500 * rb), cache_strategy=dr.CacheStrategy.LFU)
    ev_mt = _mk("mt3", 16, "cuda", storage_option=so)
    ev_ref = _mk("mt3", 16, "cuda")
    o_mt = dr.optim.AdagradOptimizer([], [ev_mt], lr=0.1, global_step=dr.optim.GlobalStep())
    o_ref = dr.optim.AdagradOptimizer([], [ev_ref], lr=0.1, global_step=dr.optim.GlobalStep())
    for step in range(12):
        lo = (step % 4) * 1500
        ids = torch.randint(lo, lo + 1500, (1200,), device="cuda")
        tgt = torch.randn(1200, 16, device="cuda")
        for ev, opt in ((ev_mt, o_mt), (ev_ref, o_ref)):
            ((ev.lookup(ids) - tgt) ** 2).sum().backward(); opt.step()
    t = ev_mt.table
    st = t.dram.tier_stats()
    assert st["ssd"]["keys"] > 0 and st["demotions"] > 0 and st["promotions"] > 0
    tiers = ev_mt.lookup_tier(torch.arange(0, 6000, device="cuda"))
    assert set(tiers.tolist()) >= {0, 1, 2}
    probe = torch.arange(0, 6000, 7, device="cuda")
    assert torch.allclose(t.lookup(probe).cpu(), ev_ref.table.lookup(probe).cpu(), atol=1e-4)
    assert ev_mt.total_count() == ev_ref.total_count()
