"""CollectiveStrategy on CPU (gloo, world_size=2): model-parallel GroupEmbedding + data-parallel dense match the single-process run."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep
    from deeprec_b200.parallel import CollectiveStrategy
    torch.manual_seed(0)
    st = CollectiveStrategy(backend="gloo")
    assert st.world_size == world and st.rank == rank
    evs = [dr.get_embedding_variable(f"mp{t}", 8, seed=11 + t) for t in range(3)]
    dense = torch.nn.Linear(24, 1)
    st.broadcast_parameters(dense)
    opt = dr.optim.AdagradOptimizer(dense.parameters(), evs, lr=0.1, global_step=GlobalStep())
    g = torch.Generator().manual_seed(100 + rank)
    for step in range(3):
        ids = [torch.randint(0, 20, (16,), generator=g) for _ in range(3)]
        sps = [dr.SparseIds.from_dense(i) for i in ids]
        with st.scope(), st.embedding_scope():
            embs = dr.group_embedding_lookup_sparse(evs, sps, ["sum"] * 3)
        y = dense(torch.cat(embs, 1)).squeeze(-1)
        loss = (y - 1.0).pow(2).sum()
        opt.zero_grad(); loss.backward()
        st.allreduce_gradients(dense.parameters())
        opt.step()
    # every table is populated ONLY on its owner
    counts = [e.total_count() for e in evs]
    for t, c in enumerate(counts):
        assert (c > 0) == (st.owner_of(t) == rank), (rank, counts)
    probe = torch.arange(20)
    rows = {t: evs[t].table.lookup(probe) for t in range(3) if st.owner_of(t) == rank}
    q.put((rank, {t: r.tolist() for t, r in rows.items()}, dense.weight.detach().tolist()))


def _single(world):
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep
    torch.manual_seed(0)
    evs = [dr.get_embedding_variable(f"sp{t}", 8, seed=11 + t) for t in range(3)]
    dense = torch.nn.Linear(24, 1)
    opt = dr.optim.AdagradOptimizer(dense.parameters(), evs, lr=0.1, global_step=GlobalStep())
    gens = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for step in range(3):
        ids = [[torch.randint(0, 20, (16,), generator=g) for _ in range(3)] for g in gens]
        cat = [torch.cat([ids[r][t] for r in range(world)]) for t in range(3)]
        embs = [e.lookup(c) for e, c in zip(evs, cat)]
        loss = (dense(torch.cat(embs, 1)).squeeze(-1) - 1.0).pow(2).sum()
        opt.zero_grad(); loss.backward(); opt.step()
    probe = torch.arange(20)
    return {t: evs[t].table.lookup(probe) for t in range(3)}, dense.weight.detach().clone()


def test_gloo_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_rows, ref_w = _single(world)
    for rank, rows, w in res:
        w = torch.tensor(w)
        assert torch.allclose(w, ref_w, atol=1e-5), (w - ref_w).abs().max()
        for t, r in rows.items():
            r = torch.tensor(r)
            assert torch.allclose(r, ref_rows[t], atol=1e-5), (t, (r - ref_rows[t]).abs().max())


def _dp_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import deeprec_b200 as dr
    from deeprec_b200.parallel import CollectiveStrategy
    st = CollectiveStrategy(backend="gloo")
    torch.manual_seed(rank)                                  # different init per rank: the estimator must broadcast rank 0's
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3), torch.nn.ReLU(), torch.nn.Linear(3, 1))
    opt = dr.optim.GradientDescentOptimizer(model, lr=0.05)
    st.scale_learning_rate(opt)
    lr_scaled = opt.param_groups[0]["lr"]
    # bucketed all-reduce with a bucket so small that the 6 parameters need several collectives, one of them a lone tensor
    for p in model.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    st.allreduce_gradients(list(model.parameters()), bucket_bytes=64)
    sums_ok = all(bool((p.grad == 3.0).all()) for p in model.parameters())
    for p in model.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    st.allreduce_gradients(list(model.parameters()), bucket_bytes=64, average=True)
    avg_ok = all(bool((p.grad == 1.5).all()) for p in model.parameters())
    g = torch.Generator().manual_seed(50 + rank)
    est = st.estimator(model, opt, lambda m, b: (m(b[0]).squeeze(-1) - b[1]).pow(2).mean(), log_every_n_steps=0)
    batches = [(torch.randn(8, 6, generator=g), torch.randn(8, generator=g)) for _ in range(4)]
    est.fit(batches, 4)
    q.put((rank, sums_ok, avg_ok, lr_scaled, [p.detach().flatten().tolist() for p in model.parameters()]))


def test_bucketed_allreduce_lr_scaling_and_estimator_keep_replicas_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res) and all(abs(r[3] - 0.1) < 1e-9 for r in res)
    assert res[0][4] == res[1][4]                            # data-parallel replicas stayed bitwise identical
