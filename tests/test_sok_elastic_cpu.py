"""SOK-style row-sharded embedding layers over gloo (world 2) vs a single-process reference, and elastic N->M re-sharding."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, tmp):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep
    from deeprec_b200.parallel import sok
    sok.Init()
    dense_l = sok.All2AllDenseEmbedding(8, slot_num=3, name="a2a")
    sparse_l = sok.DistributedEmbedding("mean", 8, name="dist")
    # same default matrix on every shard => initial values depend on the key only
    for l in (dense_l, sparse_l):
        g0 = torch.Generator().manual_seed(5)
        l.ev.default_matrix.copy_(torch.randn(l.ev.default_matrix.shape, generator=g0) * 0.1)
    evs, others = sok.split_embedding_variable_from_others(torch.nn.ModuleList([dense_l, sparse_l]))
    assert len(evs) == 2 and not others
    opt = dr.optim.AdagradOptimizer([], evs, lr=0.1, global_step=GlobalStep())
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(3):
        ids = torch.randint(0, 40, (6, 3), generator=g)
        vals = torch.randint(0, 40, (10,), generator=g)
        rows = torch.sort(torch.randint(0, 6, (10,), generator=g)).values
        w = torch.randn(6, 3, 8, generator=g); w2 = torch.randn(6, 8, generator=g)
        e1 = dense_l(ids)
        e2 = sparse_l(dr.SparseIds(vals, rows, 6))
        loss = (e1 * w).sum() + (e2 * w2).sum()
        opt.zero_grad(); loss.backward(); opt.step()
    probe = torch.arange(40)
    own = sok.row_owner(probe, world) == rank
    assert dense_l.ev.total_count() == int(((dense_l.ev.get_frequency(probe) > 0)).sum())
    out = {"dense": (probe[own].tolist(), dense_l.ev.table.lookup(probe[own]).tolist(), dense_l.ev.get_frequency(probe[own]).tolist()),
           "sparse": (probe[own].tolist(), sparse_l.ev.table.lookup(probe[own]).tolist())}
    # a key never lands on a non-owner
    assert int((dense_l.ev.get_frequency(probe[~own]) > 0).sum()) == 0
    # sok.Saver round trip with re-sharding by ownership
    sv = sok.Saver()
    sv.dump_to_file(dense_l, os.path.join(tmp, "a2a"))
    fresh = sok.All2AllDenseEmbedding(8, slot_num=3, name="a2a_restored")
    n = sv.restore_from_file(fresh, os.path.join(tmp, "a2a"))
    assert n == dense_l.ev.total_count()
    trained = probe[own][dense_l.ev.get_frequency(probe[own]) > 0]
    assert torch.allclose(fresh.ev.table.lookup(trained), dense_l.ev.table.lookup(trained))
    # sok.TFDistributedEmbedding (static table sharded id % world): forward == the full table, the gradient of a row arrives at its owner only
    full = torch.arange(50 * 4, dtype=torch.float32).view(50, 4) / 10
    tfl = sok.TFDistributedEmbedding(50, 4, initializer=lambda w: w.copy_(full[rank::world][: w.shape[0]]))
    ids_t = torch.randint(0, 50, (7, 2), generator=torch.Generator().manual_seed(7 + rank))
    e = tfl(ids_t)
    assert torch.equal(e, full[ids_t])
    (e * 2.0).sum().backward()
    all_ids = [torch.randint(0, 50, (7, 2), generator=torch.Generator().manual_seed(7 + r)) for r in range(world)]
    want = torch.zeros(50, 4); want.index_add_(0, torch.cat(all_ids).reshape(-1), torch.full((7 * 2 * world, 4), 2.0))
    assert torch.allclose(tfl.weight.grad, want[rank::world][: tfl.weight.shape[0]])
    # sok.optimizers facade
    assert type(sok.optimizers.Adam(dense_l)).__name__ == "AdamOptimizer" and type(sok.optimizers.LazyAdam([dense_l, sparse_l])).__name__ == "AdamAsyncOptimizer"
    q.put((rank, out))


def _single(world):
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep
    evd = dr.get_embedding_variable("ref_dense", 8); evs_ = dr.get_embedding_variable("ref_sparse", 8)
    for e in (evd, evs_):
        g0 = torch.Generator().manual_seed(5)
        e.default_matrix.copy_(torch.randn(e.default_matrix.shape, generator=g0) * 0.1)
    opt = dr.optim.AdagradOptimizer([], [evd, evs_], lr=0.1, global_step=GlobalStep())
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for step in range(3):
        loss = 0
        for g in gens:
            ids = torch.randint(0, 40, (6, 3), generator=g)
            vals = torch.randint(0, 40, (10,), generator=g)
            rows = torch.sort(torch.randint(0, 6, (10,), generator=g)).values
            w = torch.randn(6, 3, 8, generator=g); w2 = torch.randn(6, 8, generator=g)
            e1 = evd.lookup(ids)
            e2 = dr.embedding_lookup_sparse(evs_, dr.SparseIds(vals, rows, 6), None, "mean")
            loss = loss + (e1 * w).sum() + (e2 * w2).sum()
        opt.zero_grad(); loss.backward(); opt.step()
    return evd, evs_


def test_sok_layers_match_single_process(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, str(tmp_path))) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    evd, evs_ = _single(world)
    seen = 0
    for rank, out in res:
        keys, rows, freqs = out["dense"]
        k = torch.tensor(keys)
        assert torch.allclose(torch.tensor(rows), evd.table.lookup(k), atol=1e-5)
        assert freqs == evd.get_frequency(k).tolist()
        seen += len(keys)
        keys, rows = out["sparse"]
        assert torch.allclose(torch.tensor(rows), evs_.table.lookup(torch.tensor(keys)), atol=1e-5)
    assert seen == 40


def test_elastic_rescale_local():
    """2 shards -> 3 shards -> 1 shard: rows, optimizer slots and metadata follow their keys."""
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep
    from deeprec_b200.parallel import elastic
    shards = [dr.get_embedding_variable(f"el/part_{i}", 4, seed=3) for i in range(2)]
    opt = dr.optim.AdagradOptimizer([], shards, lr=0.1, global_step=GlobalStep())
    keys = torch.arange(0, 3000, 7)
    for _ in range(2):
        loss = 0
        own = elastic.default_owner(keys, 2)
        for i, ev in enumerate(shards):
            loss = loss + (ev.lookup(keys[own == i]) ** 2).sum()
        opt.zero_grad(); loss.backward(); opt.step()
    ref_rows = torch.zeros(keys.numel(), 4); ref_acc = torch.zeros(keys.numel(), 4)
    own = elastic.default_owner(keys, 2)
    for i, ev in enumerate(shards):
        ref_rows[own == i] = ev.table.lookup(keys[own == i]); ref_acc[own == i] = ev.table.lookup_slot(keys[own == i], 1)
    total = sum(e.total_count() for e in shards)
    third = dr.get_embedding_variable("el/part_2", 4, seed=3)
    third._set_slots(["accumulator"], [0.1], False)
    new = shards + [third]
    moved = elastic.rescale_local(shards, new)
    assert moved > 0 and sum(e.total_count() for e in new) == total
    own3 = elastic.default_owner(keys, 3)
    for i, ev in enumerate(new):
        m = own3 == i
        assert ev.total_count() == int(m.sum())
        assert torch.allclose(ev.table.lookup(keys[m]), ref_rows[m]) and torch.allclose(ev.table.lookup_slot(keys[m], 1), ref_acc[m])
        assert (ev.get_frequency(keys[m]) == 2).all()
    moved = elastic.rescale_local(new, [new[0]])
    assert new[0].total_count() == total and new[1].total_count() == 0 and new[2].total_count() == 0
    assert torch.allclose(new[0].table.lookup(keys), ref_rows)
