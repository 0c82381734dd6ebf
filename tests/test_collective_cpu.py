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
