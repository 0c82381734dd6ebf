"""CPU tier-1/2 tests of the host EmbeddingVariable engine against a dict/numpy oracle
(mirrors python/ops/embedding_variable_ops_test.py and core/kernels/embedding_variable_ops_test.cc)."""
import math

import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.optim import GlobalStep, make_optimizer


def _ev(name, dim=8, **kw):
    return dr.get_embedding_variable(name, dim, ev_option=dr.EmbeddingVariableOption(**kw), seed=3)


def test_forward_is_read_only_and_default_rows():
    ev = _ev("ro")
    ids = torch.tensor([5, 5 + 4096, 7])
    e = ev.lookup(ids)
    assert ev.total_count() == 0                      # forward never inserts (embedding_var.h:202-219)
    assert torch.equal(e[0], e[1])                    # default row = key % default_value_dim
    assert torch.equal(e[0], ev.default_matrix[5])


def test_adagrad_matches_manual_math_with_dedup_and_counts():
    ev = _ev("ada")
    opt = dr.optim.AdagradOptimizer([], [ev], lr=0.1, initial_accumulator_value=0.1, global_step=GlobalStep())
    ids = torch.tensor([1, 2, 2, 2])
    w0 = ev.lookup(torch.tensor([1, 2])).detach().clone()
    e = ev.lookup(ids)
    (e * torch.tensor([[1.0], [2.0], [3.0], [4.0]])).sum().backward()
    opt.step()
    g = torch.tensor([1.0, 9.0])                      # per-key summed gradient (dedup before the accumulator update)
    acc = 0.1 + g ** 2
    expect = w0 - 0.1 * g.unsqueeze(1) / acc.sqrt().unsqueeze(1)
    got = ev.lookup(torch.tensor([1, 2])).detach()
    assert torch.allclose(got, expect, atol=1e-6)
    assert ev.get_frequency(torch.tensor([1, 2, 3])).tolist() == [1, 3, 0]
    assert ev.get_version(torch.tensor([1, 3])).tolist() == [0, -1]
    assert torch.allclose(ev.slot_values(torch.tensor([2]), "accumulator"), acc[1].expand(1, 8))


@pytest.mark.parametrize("name", ["adagrad", "adagraddecay", "adam", "adamasync", "adamw", "ftrl", "gradientdescent"])
def test_sparse_rules_equal_dense_rules_on_a_dense_variable(name):
    """Every id touched every step => the sparse EV update must equal the dense-parameter update of the
    same optimizer (the reference tests compare EV training against a plain tf.Variable model)."""
    torch.manual_seed(0)
    n, dim = 6, 8
    ev = _ev(f"eq_{name}", dim)
    init = ev.lookup(torch.arange(n)).detach().clone()
    dense = torch.nn.Parameter(init.clone())
    kw = dict(lr=0.05)
    if name == "adagraddecay":
        kw.update(accumulator_decay_step=2, accumulator_decay_rate=0.5)
    gs1, gs2 = GlobalStep(), GlobalStep()
    o_ev = make_optimizer(name, [], [ev], global_step=gs1, **kw)
    o_dn = make_optimizer(name, [dense], None, global_step=gs2, **kw)
    for step in range(5):
        tgt = torch.randn(n, dim)
        ((ev.lookup(torch.arange(n)) - tgt) ** 2).sum().backward(); o_ev.step()
        o_dn.zero_grad(); ((dense - tgt) ** 2).sum().backward(); o_dn.step()
    got = ev.lookup(torch.arange(n)).detach()
    if name == "ftrl":
        # EV FTRL is the row-norm (group-lasso) form with l1 = 0 => coef*linear, dense is element-wise: both reduce to
        # -linear/quadratic when l1 == 0
        pass
    assert torch.allclose(got, dense.detach(), atol=2e-5, rtol=1e-4), (got - dense).abs().max()


def test_counter_filter_admission_and_no_permission_value():
    ev = _ev("cf", filter_option=dr.CounterFilter(3), init_option=dr.InitializerOption(default_value_no_permission=0.5))
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=1.0, global_step=GlobalStep())
    ids = torch.tensor([9])
    for step in range(2):
        e = ev.lookup(ids)
        assert torch.all(e == 0.5)                    # not admitted yet
        e.sum().backward(); opt.step()
    assert ev.total_count() == 0 and ev.table.total_keys() == 1
    e = ev.lookup(ids); e.sum().backward(); opt.step()    # third occurrence: admitted inside the apply, grad applied
    assert ev.total_count() == 1
    assert torch.allclose(ev.lookup(ids).detach(), ev.default_matrix[9].unsqueeze(0) - 1.0)


def test_bloom_filter_admission():
    ev = _ev("cbf", filter_option=dr.CBFFilter(filter_freq=3, max_element_size=1000, false_positive_probability=0.01,
                                               counter_type=torch.int16))
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=1.0, global_step=GlobalStep())
    for step in range(2):
        ev.lookup(torch.tensor([4])).sum().backward(); opt.step()
        assert ev.total_count() == 0 and ev.table.total_keys() == 0      # bloom keeps no per-key state
    assert ev.get_frequency(torch.tensor([4])).item() >= 2
    ev.lookup(torch.tensor([4])).sum().backward(); opt.step()
    assert ev.total_count() == 1
    k, m = math.ceil(math.log2(1 / 0.01)), math.ceil(1000 * abs(math.log(0.01)) / math.log(2) ** 2)
    assert ev.table.bloom_state().numel() == m * 2 and k == 7


def test_global_step_and_l2_eviction_only_at_shrink():
    ev = _ev("gs", evict_option=dr.GlobalStepEvict(steps_to_live=2))
    gs = GlobalStep()
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=0.1, global_step=gs)
    ev.lookup(torch.tensor([1, 2])).sum().backward(); opt.step()          # version 0
    for _ in range(4):
        ev.lookup(torch.tensor([2])).sum().backward(); opt.step()         # key 2 stays fresh
    assert ev.total_count() == 2                                          # nothing evicted during training
    assert ev.table.shrink(int(gs)) == 1
    assert ev.total_count() == 1 and ev.get_version(torch.tensor([1])).item() == -1
    ev2 = _ev("l2", evict_option=dr.L2WeightEvict(l2_weight_threshold=1e9))
    o2 = dr.optim.GradientDescentOptimizer([], [ev2], lr=0.1, global_step=GlobalStep())
    ev2.lookup(torch.tensor([1, 2, 3])).sum().backward(); o2.step()
    assert ev2.table.shrink(0) == 3 and ev2.total_count() == 0


def test_snapshot_bucket_order_and_resharded_import():
    ev = _ev("snap")
    opt = dr.optim.AdagradOptimizer([], [ev], lr=0.1, global_step=GlobalStep())
    ids = torch.arange(0, 5000, 7)
    ev.lookup(ids).sum().backward(); opt.step()
    s = ev.table.snapshot()
    b = s["keys"] % 1000
    assert torch.all(b[1:] >= b[:-1]) and int(s["partition_offset"][-1]) == ids.numel()
    for bk in (0, 13, 999):
        lo, hi = int(s["partition_offset"][bk]), int(s["partition_offset"][bk + 1])
        assert torch.all(s["keys"][lo:hi] % 1000 == bk)
    # N -> M re-shard: 3 partitions keep key % 1000 % 3 == p, union is everything
    total = 0
    for p in range(3):
        part = _ev(f"snap_p{p}")
        part._set_slots(["accumulator"], [0.1], False)
        kept = part.table.import_(s["keys"], s["rows"], s["freqs"], s["versions"], p, 3)
        total += kept
        pk = part.table.snapshot()["keys"]
        assert torch.all(pk % 1000 % 3 == p)
        assert torch.allclose(part.table.lookup(pk), ev.table.lookup(pk))
    assert total == ids.numel()


def test_incremental_dirty_tracking():
    ev = _ev("dirty")
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=0.1, global_step=GlobalStep())
    ev.lookup(torch.arange(10)).sum().backward(); opt.step()
    ev.table.clear_dirty()
    ev.lookup(torch.tensor([3, 4])).sum().backward(); opt.step()
    d = ev.table.snapshot(dirty_only=True)
    assert sorted(d["keys"].tolist()) == [3, 4]


def test_concurrent_growth_many_keys():
    ev = _ev("big", dim=4, init_capacity=64)
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=1.0, global_step=GlobalStep())
    ids = torch.randperm(300000)
    ev.lookup(ids).sum().backward(); opt.step()
    assert ev.total_count() == 300000
    probe = ids[:1000]
    assert torch.allclose(ev.lookup(probe).detach(), ev.default_matrix[probe % 4096] - 1.0)
    assert ev.table.remove(probe[:10]) == 10 and ev.total_count() == 299990


def test_partitioned_multihash_dynamic_variants():
    pev = dr.get_embedding_variable("part", 8, partitioner=dr.fixed_size_partitioner(4))
    opt = dr.optim.AdagradOptimizer(pev, None, lr=0.1, global_step=GlobalStep())
    ids = torch.arange(40)
    pev.lookup(ids).sum().backward(); opt.step()
    assert pev.total_count() == 40 and [p.total_count() for p in pev.parts] == [10] * 4
    mh = dr.get_multihash_variable("mh", [[10, 8], [7, 8]], operation="add")
    assert mh.lookup(torch.tensor([0, 69])).shape == (2, 8)
    assert dr.get_multihash_variable("mh2", [[10, 4], [7, 6]], operation="concat").lookup(torch.tensor([3])).shape == (1, 10)
    dv = dr.get_dynamic_dimension_embedding_variable("dyn", 4, 3)
    out = dv.lookup(torch.tensor([1, 2]), torch.tensor([1, 3]))
    assert out.shape == (2, 12) and torch.all(out[0, 4:] == 0) and out[1, 8:].abs().sum() > 0


def test_dram_ssdhash_tier_matches_dram_only(tmp_path):
    """DRAM_SSDHASH: a DRAM tier of ~200 rows over the log-structured SSD store trains exactly like an all-DRAM table."""
    import deeprec_b200 as dr
    dim = 8
    row_bytes = 4 * dim * 2                      # emb + adagrad accumulator
    so = dr.StorageOption(dr.StorageType.DRAM_SSDHASH, storage_path=str(tmp_path), storage_size=(200 * row_bytes,))
    ev_t = dr.get_embedding_variable("ssd_tiered", dim, ev_option=dr.EmbeddingVariableOption(storage_option=so), seed=11)
    ev_r = dr.get_embedding_variable("ssd_ref", dim, seed=11)
    opt_t, opt_r = dr.optim.AdagradOptimizer([], [ev_t], lr=0.1), dr.optim.AdagradOptimizer([], [ev_r], lr=0.1)
    g = torch.Generator().manual_seed(0)
    for step in range(12):
        ids = torch.randint(0, 1500, (256,), generator=g)
        w = torch.randn(256, dim, generator=g)
        for ev, opt in ((ev_t, opt_t), (ev_r, opt_r)):
            opt.zero_grad()
            (ev.lookup(ids) * w).sum().backward()
            opt.step()
    t = ev_t.table
    st = t.tier_stats()
    assert st["dram_rows"] <= 200 + 64 and st["ssd"]["keys"] > 0 and st["demotions"] > 0 and st["promotions"] > 0
    assert ev_t.total_count() == ev_r.total_count()
    probe = torch.arange(0, 1500)
    tiers = t.lookup_tier(probe)
    assert (tiers == 1).any() and (tiers == 0).any()
    assert torch.equal(ev_t.get_frequency(probe), ev_r.get_frequency(probe))
    assert torch.allclose(ev_t.lookup(probe).detach(), ev_r.lookup(probe).detach(), atol=1e-6)
    # checkpoint covers both tiers
    snap_t, snap_r = t.snapshot(), ev_r.table.snapshot()
    assert torch.equal(snap_t["keys"], snap_r["keys"]) and torch.allclose(snap_t["rows"], snap_r["rows"], atol=1e-6)
    # overwrite churn -> dead records -> compaction reclaims files
    ssd = t.ssd
    ks = ssd.keys()[:64]
    rows, f, v, found = ssd.get(ks)
    assert found.all()
    before = ssd.stats()
    for _ in range(40):
        ssd.put(ks, rows, f, v)
    ssd.compact(0.0)
    after = ssd.stats()
    assert after["keys"] == before["keys"] and after["bytes"] <= before["bytes"] + 64 * (24 + 4 * t.stride) * 2
    rows2, _, _, found2 = ssd.get(ks)
    assert found2.all() and torch.equal(rows2, rows)


def test_ssd_store_async_compaction(tmp_path):
    from deeprec_b200.ops.host_tiers import SsdStore
    s = SsdStore(4, str(tmp_path), file_bytes=40 * 200, async_compaction=True)     # 200 records per file
    keys = torch.arange(1000)
    for rep in range(6):
        s.put(keys, torch.full((1000, 4), float(rep)), torch.full((1000,), rep), torch.full((1000,), rep))
    import time
    for _ in range(100):
        if s.stats()["compactions"] > 0 and s.stats()["files"] <= 12:
            break
        time.sleep(0.05)
    st = s.stats()
    assert st["keys"] == 1000 and st["compactions"] > 0
    rows, f, v, found = s.get(keys)
    assert found.all() and (rows == 5.0).all() and (f == 5).all()
    assert s.remove(keys[:500]) == 500 and s.size() == 500


def test_grouped_host_lookup_and_apply_equal_per_table_path():
    """ops/host_group: one native lookup -> [B, T, D] and one native dedup+apply for T tables == T independent lookups + applies,
    including admission filters, frequency / version bookkeeping and default rows."""
    from deeprec_b200.ops.host_group import group_lookup_dense_host
    T, B, D = 9, 300, 8        # T >= pool size on small boxes exercises the table-parallel branch, see also T=3 below

    def make(tag, T):
        dr.embedding_variable.clear_registry()
        evs = [dr.get_embedding_variable(f"{tag}/t{t}", D, seed=20 + t,
                                         ev_option=dr.EmbeddingVariableOption(filter_option=dr.CounterFilter(2) if t % 2 else None)) for t in range(T)]
        return evs, dr.optim.AdamOptimizer([], evs, lr=0.01, global_step=GlobalStep())

    for T_ in (T, 3):
        ga, oa = make("grp", T_)
        gb, ob = make("ref", T_)
        g = torch.Generator().manual_seed(5)
        for step in range(4):
            ids = (torch.randn(T_, B, generator=g).abs() * 40).long()
            w = torch.randn(B, T_, D, generator=g)
            out_a = group_lookup_dense_host(ga, ids)
            out_b = torch.stack([e.lookup(ids[t]) for t, e in enumerate(gb)], 1)
            assert out_a.shape == (B, T_, D) and torch.equal(out_a, out_b)
            (out_a * w).sum().backward(); oa.step(); oa.zero_grad()
            (out_b * w).sum().backward(); ob.step(); ob.zero_grad()
        probe = torch.arange(0, 200)
        for ea, eb in zip(ga, gb):
            assert ea.total_count() == eb.total_count() and ea.table.total_keys() == eb.table.total_keys()
            assert torch.allclose(ea.table.lookup(probe), eb.table.lookup(probe), atol=1e-6)
            assert torch.equal(ea.get_frequency(probe), eb.get_frequency(probe)) and torch.equal(ea.get_version(probe), eb.get_version(probe))
            assert torch.allclose(ea.slot_values(probe, "v"), eb.slot_values(probe, "v"), atol=1e-7)
        with torch.no_grad():                                      # inference path: same native call, no autograd node
            assert torch.equal(group_lookup_dense_host(ga, ids), torch.stack([e.lookup(ids[t]) for t, e in enumerate(ga)], 1))
    # not eligible -> None (mixed dims)
    dr.embedding_variable.clear_registry()
    mixed = [dr.get_embedding_variable("m0", 8), dr.get_embedding_variable("m1", 4)]
    assert group_lookup_dense_host(mixed, torch.zeros(2, 5, dtype=torch.int64)) is None


def test_pad_key_reads_zeros_and_is_never_created_or_counted():
    from deeprec_b200.config import PAD_KEY
    dr.embedding_variable.clear_registry()
    ev = _ev("padkey", 4, filter_option=dr.CounterFilter(1))
    opt = dr.optim.AdagradOptimizer([], [ev], lr=0.5, global_step=GlobalStep())
    ids = torch.tensor([[5, PAD_KEY, 7], [PAD_KEY, PAD_KEY, 5]])
    for _ in range(2):
        out = ev.lookup(ids)
        assert torch.all(out[0, 1] == 0) and torch.all(out[1, :2] == 0)
        (out * torch.arange(1.0, 7.0).view(2, 3, 1)).sum().backward(); opt.step(); opt.zero_grad()
    assert ev.total_count() == 2 and ev.table.total_keys() == 2
    assert ev.get_frequency(torch.tensor([5, 7, PAD_KEY])).tolist() == [4, 2, 0] and ev.get_version(torch.tensor([PAD_KEY])).tolist() == [-1]
    assert torch.all(ev.lookup(torch.tensor([PAD_KEY])) == 0) and torch.all(ev.slot_values(torch.tensor([PAD_KEY]), "accumulator") == 0.1)
    keys = ev.export()[0]
    assert sorted(keys.tolist()) == [5, 7]


@pytest.mark.parametrize("n_ids,L", [(300, 7), (2100, 12)])        # the larger case crosses into the parallel bucketed dedup (>= 16384 ids)
def test_pooled_lookup_and_multi_segment_apply_equal_the_materialised_path(n_ids, L):
    """lookup_pooled (no [B, L, D] intermediate, one gradient row per bag) mixed with a plain lookup of the SAME table in one step
    == lookup + mask + sum with every per-occurrence gradient materialised: values, frequencies, versions, optimizer slots."""
    from deeprec_b200.config import PAD_KEY
    B, D = n_ids, 8
    dr.embedding_variable.clear_registry()
    a = _ev(f"pool_a{L}", D, filter_option=dr.CounterFilter(2)); b = _ev(f"pool_b{L}", D, filter_option=dr.CounterFilter(2))
    b.default_matrix.copy_(a.default_matrix); b._table = None
    oa = dr.optim.AdagradOptimizer([], [a], lr=0.1, global_step=GlobalStep()); ob = dr.optim.AdagradOptimizer([], [b], lr=0.1, global_step=GlobalStep())
    g = torch.Generator().manual_seed(9)
    for step in range(3):
        hist = (torch.randn(B, L, generator=g).abs() * 30).long()
        lens = torch.randint(0, L + 1, (B,), generator=g)
        mask = torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)
        tgt = torch.randint(0, 60, (B,), generator=g)
        w1, w2 = torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
        pa = a.lookup_pooled(hist, mask)
        ((pa * w1).sum() + (a.lookup(tgt) * w2).sum()).backward(); oa.step(); oa.zero_grad()
        rows = b.lookup(torch.where(mask, hist, torch.zeros_like(hist)))              # reference: materialise, mask, sum
        pb = (rows * mask.unsqueeze(-1)).sum(1)
        assert torch.allclose(pa.detach(), pb.detach(), atol=1e-5)
        # feed the reference exactly the valid occurrences (the padded positions must not touch key 0)
        b._pending.clear()
        ((b.lookup(hist[mask]) * w1.unsqueeze(1).expand(B, L, D)[mask]).sum() + (b.lookup(tgt) * w2).sum()).backward(); ob.step(); ob.zero_grad()
    probe = torch.arange(0, 150)
    assert a.total_count() == b.total_count() and a.table.total_keys() == b.table.total_keys()
    assert torch.allclose(a.table.lookup(probe), b.table.lookup(probe), atol=1e-5)
    assert torch.equal(a.get_frequency(probe), b.get_frequency(probe)) and torch.equal(a.get_version(probe), b.get_version(probe))
    assert torch.allclose(a.slot_values(probe, "accumulator"), b.slot_values(probe, "accumulator"), atol=1e-5)
    with torch.no_grad():
        assert torch.allclose(a.lookup_pooled(torch.where(mask, hist, PAD_KEY)), pb.detach() * 0 + a.lookup_pooled(hist, mask))
