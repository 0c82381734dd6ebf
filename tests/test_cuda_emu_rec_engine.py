"""FusedRecEngine on the CUDA-on-CPU emulation (``_native.cuda_emulation()``): the whole training step of the generic engine -- unique-first sparse
pipeline, autograd dense net, fused dense + row-wise optimizers, multi-tier tables, checkpoints -- runs in the CPU test tier, eagerly (no CUDA
graph), against the same oracles as tests/test_gpu_rec_engine.py; with the ranks as THREADS (parallel/emu_comm.py) the model-parallel step incl.
the fused all-reduce + optimizer kernel and owner-side tier promotion runs rank against rank.  GPU twins: test_gpu_rec_engine.py, test_gpu_tier.py,
test_gpu_multi.py."""
import copy
import threading

import pytest
import torch

from deeprec_b200 import _native

pytestmark = [pytest.mark.timeout(1200)]
CARDS = [50, 1000, 7, 300] + [97] * 22


def _batch(B, seed, lo=None):
    g = torch.Generator().manual_seed(seed)
    ids = torch.stack([torch.randint(0, c, (B,), generator=g) for c in CARDS])
    if lo is not None:
        ids[1] = torch.randint(lo, lo + 250, (B,), generator=g)
    return ids, (torch.rand(B, generator=g) < 0.3).float(), {"dense": torch.rand(B, 13, generator=g) * 3}


def test_deepfm_engine_on_the_emulation_matches_a_torch_replica():
    import deeprec_b200 as dr
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    B = 256
    with _native.cuda_emulation():
        model = build_model("deepfm", device="cpu")
        eng = criteo_engine(model, B, table_rows=CARDS, learning_rate=0.05)
        ref_net = copy.deepcopy(eng.net.inner)
        tables = torch.nn.ModuleList([torch.nn.Embedding(c, 16) for c in CARDS])
        with torch.no_grad():
            for t, emb in enumerate(tables):
                dm = eng.tables[t].default_matrix
                emb.weight.copy_(dm[torch.arange(CARDS[t]) % dm.shape[0]])
        opt = torch.optim.Adagrad(list(ref_net.parameters()) + list(tables.parameters()), lr=0.05, initial_accumulator_value=0.1, eps=0.0)
        le, lr_ = [], []
        for s in range(5):
            ids, y, dense = _batch(B, 100 + s)
            eng.load_batch(ids, y, dense)
            eng.train_step()
            le.append(eng.loss_value())
            embs = torch.stack([tables[t](ids[t]) for t in range(26)], 1)
            loss = torch.nn.functional.binary_cross_entropy_with_logits(ref_net.logits(dense["dense"], embs).float(), y)
            opt.zero_grad(); loss.backward(); opt.step()
            lr_.append(loss.item())
        for a, b in zip(le, lr_):
            assert abs(a - b) < 0.03 * max(1.0, abs(b)), (le, lr_)
        keys = ids[1][:64]
        assert (eng.tables[1].lookup(keys) - tables[1].weight[keys]).abs().max().item() < 0.05
        assert eng.tables[0].overflowed() == 0
        p = eng.predict()
        assert p.shape == (B,) and bool(((p >= 0) & (p <= 1)).all())


def test_engine_checkpoint_and_tiered_tables_on_the_emulation(tmp_path):
    """(a) a 256-row HBM cache over the DRAM tier trains like single-tier tables; (b) save (full + incremental, both tiers) -> restore into a fresh
    engine -> identical continuation."""
    import deeprec_b200 as dr
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    B = 128
    batches = [_batch(B, 200 + s, lo=(s % 4) * 250) for s in range(10)]              # table 1: 1000 distinct ids rotate through a 256-row cache
    with _native.cuda_emulation():
        engs = []
        for tiered in (None, {1: {"cache_rows": 256, "strategy": 0}}, {1: {"cache_rows": 256, "strategy": 0}}):
            dr.embedding_variable.clear_registry()
            torch.manual_seed(0)
            engs.append(criteo_engine(build_model("deepfm", device="cpu"), B, table_rows=CARDS, learning_rate=0.05, tiered=tiered))
        ref, tier, fresh = engs
        la, lb = [], []
        tier.prefetch(batches[0][0])
        for s in range(7):
            for e, l in ((ref, la), (tier, lb)):
                e.load_batch(*batches[s]); e.train_step(); l.append(e.loss_value())
            tier.prefetch(batches[s + 1][0])
            if s == 3:
                tier.save(str(tmp_path / "ck"))
        assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-3, (la, lb)
        st = tier.tiers[1][0].stats()
        assert st["demoted_rows"] > 0 and st["promoted_rows"] > 0, st
        tier.save(str(tmp_path / "ck"), incremental=True)
        assert fresh.restore(str(tmp_path / "ck")) == 7
        assert torch.equal(fresh.params, tier.params)
        probe = torch.arange(0, 1000, 3)
        assert torch.allclose(fresh.tiers[1][0].lookup(probe), tier.tiers[1][0].lookup(probe), atol=1e-6)
        lc, ld = [], []
        fresh.prefetch(batches[7][0])
        for s in range(7, 10):
            for e, l in ((tier, lc), (fresh, ld)):
                e.load_batch(*batches[s]); e.train_step(); l.append(e.loss_value())
            if s + 1 < 10:
                tier.prefetch(batches[s + 1][0]); fresh.prefetch(batches[s + 1][0])
        assert max(abs(a - b) for a, b in zip(lc, ld)) < 2e-3, (lc, ld)


def _rank(rank, W, shared, B, steps, tiered, out, errors):
    try:
        import deeprec_b200 as dr
        from deeprec_b200.models.rec_engine import criteo_engine
        from deeprec_b200.models.zoo import build_model
        from deeprec_b200.parallel.emu_comm import EmuComm
        with _native.cuda_emulation():
            with shared.lock:                                    # the default RNG and the registry of framework-API variables are process-global: one model at a time
                torch.manual_seed(0)
                dr.embedding_variable.clear_registry()
                model = build_model("deepfm", device="cpu")
            comm = EmuComm(shared, rank) if W > 1 else None
            eng = criteo_engine(model, B, table_rows=CARDS, learning_rate=0.05, rank=rank, world_size=W, comm=comm, tiered=tiered)
            losses = []
            nxt = lambda s: _batch(B, 300 + 10 * s + rank, lo=(s % 4) * 250)
            if tiered:
                eng.prefetch(nxt(0)[0])
            for s in range(steps):
                eng.load_batch(*nxt(s)); eng.train_step(); losses.append(eng.loss_value())
                if tiered and s + 1 < steps:
                    eng.prefetch(nxt(s + 1)[0])
            probe = torch.arange(0, 1000)
            rows = eng.tiers[1][0].lookup(probe) if tiered else eng.tables[1].lookup(probe)
            out[rank] = (losses, eng.params.clone(), rows.clone())
            if comm is not None:
                comm.host_barrier()
    except BaseException as e:                                  # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            shared.barrier.abort()
        except Exception:
            pass
        raise


def _run_world(W, B, steps, tiered=None):
    from deeprec_b200.parallel.emu_comm import EmuWorld
    shared, out, errors = EmuWorld(W), {}, []
    threads = [threading.Thread(target=_rank, args=(r, W, shared, B, steps, tiered, out, errors)) for r in range(W)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1000)
    assert not errors and len(out) == W, errors
    return out


def test_model_parallel_engine_ranks_as_threads():
    """World 2 (per-rank batch B) == world 1 on the concatenated batch is not testable with BatchNorm in the net, so the checks are the invariants of
    the data-parallel dense part + model-parallel tables: identical dense replicas after every fused all-reduce + optimizer step, identical global loss
    on both ranks, falling loss; and tiered tables (owner-side promotion over peer memory) == single-tier tables, rank by rank."""
    from deeprec_b200.checkpoint.engine_ckpt import sp_owner
    with _native.cuda_emulation():
        pass
    W, B, steps = 2, 128, 6
    plain = _run_world(W, B, steps)
    assert torch.equal(plain[0][1], plain[1][1]), "dense replicas diverged"
    assert max(abs(a - b) for a, b in zip(plain[0][0], plain[1][0])) < 1e-6
    assert all(l == l for l in plain[0][0])
    tier = _run_world(W, B, steps, tiered={1: {"cache_rows": 128, "strategy": 0}})
    assert max(abs(a - b) for a, b in zip(plain[0][0], tier[0][0])) < 2e-3, (plain[0][0], tier[0][0])
    owner = sp_owner(torch.arange(0, 1000), W)
    for r in range(W):
        mine = owner == r
        assert torch.allclose(plain[r][2][mine], tier[r][2][mine], atol=1e-4), r


def test_micro_batches_through_the_dense_net_equal_one_big_batch():
    """FusedRecEngine(micro_batch_num=4) (auto micro-batch, graph_execution_state.cc:635-729): a net without batch statistics must train exactly as
    with the whole batch at once -- slice losses average to the batch loss, parameter gradients accumulate, embedding gradients land in their slices."""
    from deeprec_b200.models.rec_engine import FusedRecEngine
    B, C, D = 128, 4, 16
    res = []
    with _native.cuda_emulation():
        for M in (1, 4):
            torch.manual_seed(3)
            net = torch.nn.Sequential(torch.nn.Linear(C * D + 5, 32), torch.nn.ReLU(), torch.nn.Linear(32, 1))
            fwd = lambda net, dense, emb, ids: net(torch.cat([emb.float().flatten(1), dense["x"]], 1)).squeeze(1)
            eng = FusedRecEngine(net, fwd, [0, 1, 1, 2], [50, 300, 20], B, embedding_dim=D, dense_inputs={"x": ((B, 5), torch.float32)}, learning_rate=0.05,
                                 micro_batch_num=M)
            losses = []
            for s in range(5):
                g = torch.Generator().manual_seed(50 + s)
                ids = torch.stack([torch.randint(0, 50, (B,), generator=g), torch.randint(0, 300, (B,), generator=g), torch.randint(0, 300, (B,), generator=g),
                                   torch.randint(0, 20, (B,), generator=g)])
                eng.load_batch(ids, (torch.rand(B, generator=g) < 0.4).float(), {"x": torch.rand(B, 5, generator=g)})
                eng.train_step(); losses.append(eng.loss_value())
            res.append((losses, eng.params.clone(), eng.tables[1].lookup(torch.arange(300)).clone()))
    assert max(abs(a - b) for a, b in zip(res[0][0], res[1][0])) < 1e-5, (res[0][0], res[1][0])
    assert torch.allclose(res[0][1], res[1][1], atol=1e-5) and torch.allclose(res[0][2], res[1][2], atol=1e-4)
    with pytest.raises(ValueError):
        with _native.cuda_emulation():
            FusedRecEngine(torch.nn.Linear(16, 1), lambda *a: None, [0], [10], 10, micro_batch_num=3)
