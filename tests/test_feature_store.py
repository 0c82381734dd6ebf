"""Remote feature store: native RESP client against the in-process mini Redis server; serving a zoo model out of the store."""
import threading

import numpy as np
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.models.zoo import build_model
from deeprec_b200.serving.feature_store import (LocalFeatureStore, MiniRedisServer, RedisFeatureStore, attach_feature_store,
                                                export_delta_to_feature_store, export_to_feature_store)


def test_redis_client_rows_meta_and_pipelining():
    with MiniRedisServer() as srv:
        st = RedisFeatureStore(srv.host, srv.port, pool_size=2)
        assert st.ping() and st.dbsize() == 0
        rng = np.random.default_rng(0)
        keys = rng.choice(1 << 40, size=3000, replace=False).astype(np.int64) - (1 << 20)      # > 512 keys -> several pipelined commands
        rows = rng.standard_normal((3000, 16)).astype(np.float32)
        before = srv.commands
        st.insert("m/1/C1", keys, rows)
        assert srv.commands - before == 6 and st.dbsize() == 3000
        q = np.concatenate([keys[::-1][:1500], np.array([-7, 123456789], np.int64)])
        got, found = st.lookup("m/1/C1", q, 16)
        assert found[:1500].all() and not found[1500:].any() and np.array_equal(got[:1500], rows[::-1][:1500])
        assert not st.lookup("m/1/C2", keys[:4], 16)[1].any()                                # other table: different key space
        assert not st.lookup("m/1/C1", keys[:4], 8)[1].any()                                 # wrong dim is a miss, not garbage
        assert st.remove("m/1/C1", keys[:10]) == 10 and st.dbsize() == 2990
        st.set_meta("m/latest_version", "17"); assert st.get_meta("m/latest_version") == "17" and st.get_meta("nope") is None
        # concurrent callers share the pool
        errs = []
        def worker(i):
            try:
                g, f = st.lookup("m/1/C1", keys[100 + i * 50: 150 + i * 50], 16)
                assert f.all() and np.array_equal(g, rows[100 + i * 50: 150 + i * 50])
            except Exception as e:      # pragma: no cover
                errs.append(e)
        ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
        [t.start() for t in ts]; [t.join() for t in ts]
        assert not errs
        st.flush(); assert st.dbsize() == 0
        st.close()


def test_redis_auth_and_connection_errors():
    with MiniRedisServer(password="s3cret") as srv:
        st = RedisFeatureStore(srv.host, srv.port, password="s3cret", pool_size=1)
        assert st.ping()
        st.close()
        with pytest.raises(ConnectionError, match="invalid password"):
            RedisFeatureStore(srv.host, srv.port, password="wrong", pool_size=1)
        port = srv.port
    with pytest.raises(ConnectionError):
        RedisFeatureStore("127.0.0.1", port, pool_size=1, timeout_ms=500)


@pytest.mark.parametrize("kind", ["local", "redis"])
def test_model_served_from_feature_store_matches_local(kind):
    torch.manual_seed(0)
    dr.embedding_variable.clear_registry()
    model = build_model("dlrm", device="cpu")
    opt = dr.optim.AdagradOptimizer(model, lr=0.05)
    g = torch.Generator().manual_seed(1)
    dense = torch.randn(64, 13, generator=g); ids = torch.randint(0, 200, (26, 64), generator=g); y = (torch.rand(64, generator=g) < 0.4).float()
    for _ in range(3):
        opt.zero_grad(); model.loss(dense, ids, y).backward(); opt.step()
    model.eval()
    probe = torch.randint(0, 400, (26, 64), generator=g)                  # half of these ids were never trained -> default rows
    with torch.no_grad():
        ref = model(dense, probe).clone()
    from deeprec_b200.optim.optimizers import collect_embedding_variables
    evs = collect_embedding_variables(model)
    srv = MiniRedisServer() if kind == "redis" else None
    store = RedisFeatureStore(srv.host, srv.port) if srv else LocalFeatureStore()
    try:
        n = export_to_feature_store(evs, store, "dlrm", version=3)
        assert n == sum(ev.total_count() for ev in evs) and store.get_meta("dlrm/latest_version") == "3"
        # delta: two more steps touch a few rows; only those are re-sent
        for ev in evs:
            ev.table.clear_dirty()
        model.train()
        for _ in range(2):
            opt.zero_grad(); model.loss(dense[:8], ids[:, :8], y[:8]).backward(); opt.step()
        model.eval()
        nd = export_delta_to_feature_store(evs, store, "dlrm", version=3)
        assert 0 < nd < n
        with torch.no_grad():
            ref2 = model(dense, probe).clone()
        swapped = attach_feature_store(model, store, "dlrm")
        assert len(swapped) == 26 and not any(isinstance(t, dr.EmbeddingVariable) for t in model.modules())
        with torch.no_grad():
            got = model(dense, probe)
        assert torch.allclose(got, ref2, atol=1e-6) and not torch.allclose(ref, ref2, atol=1e-6)
    finally:
        store.close()
        if srv:
            srv.close()
