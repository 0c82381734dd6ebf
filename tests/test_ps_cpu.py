"""Asynchronous parameter-server mode on CPU: 2 PS + 2 workers over torch.distributed.rpc."""
import json
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


NUM_PS, NUM_WORKERS, STEPS = 2, 2, 6


def _spawn_roles(target, n_roles, tmp, timeout=240, attempts=2):
    """Start one process per role and wait for all of them.  The rendezvous goes through a TCP port picked just before the processes start
    and through torch.distributed.rpc's own timeouts: on a box that is busy with other work (a compiler, another test run) either can fail
    without anything being wrong with the code under test, so a failed ROLE START is retried once on a fresh port; assertion failures of the
    test body are never retried (they happen after this returns)."""
    import shutil
    ctx = mp.get_context("spawn")
    for attempt in range(attempts):
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, port, tmp)) for r in range(n_roles)]
        for p in procs:
            p.start()
        try:
            _join_all(procs, timeout)
            return
        except AssertionError:
            if attempt + 1 == attempts:
                raise
            for f in os.listdir(tmp):                          # partial outputs of the failed attempt
                q = os.path.join(tmp, f)
                shutil.rmtree(q, ignore_errors=True) if os.path.isdir(q) else os.remove(q)


def _join_all(procs, timeout=240):
    """Wait for every process; if one fails the rest (blocked in rpc.shutdown) are terminated instead of hanging the test."""
    import time
    t0 = time.time()
    while time.time() - t0 < timeout:
        codes = [p.exitcode for p in procs]
        if all(c is not None for c in codes) or any(c not in (None, 0) for c in codes):
            break
        time.sleep(0.2)
    codes = [p.exitcode for p in procs]
    for p in procs:
        if p.exitcode is None:
            p.terminate()
            p.join(10)
    assert all(c == 0 for c in codes), f"PS/worker exit codes {codes}"


def _proc(rank, port, tmp):
    torch.manual_seed(0)
    from deeprec_b200.parallel import ps
    if rank < NUM_PS:
        ps.run_ps(rank, NUM_PS, NUM_WORKERS, port, stats_path=os.path.join(tmp, f"ps{rank}.json"))
        return
    j = rank - NUM_PS
    client = ps.PSClient(j, NUM_PS, NUM_WORKERS, port, slice_bytes=64)       # tiny slices: the dense transfer is really sliced
    user = client.create_embedding("user", 8, optimizer="adagrad", lr=0.1, seed=3)
    item = client.create_embedding("item", 8, optimizer="adagrad", lr=0.1, seed=4)
    dense = torch.nn.Linear(16, 1)
    client.register_dense(dense, lr=0.05)
    w0 = dense.weight.detach().clone()
    g = torch.Generator().manual_seed(10 + j)
    losses = []
    for step in range(STEPS):
        client.pull_dense()
        uid = torch.randint(0, 50, (32,), generator=g); iid = torch.randint(0, 80, (32,), generator=g)
        u, it = ps.group_pull(client, [user, item], [uid, iid])               # ONE rpc per PS for both tables
        y = ((uid + iid) % 2).float()
        loss = torch.nn.functional.binary_cross_entropy_with_logits(dense(torch.cat([u, it], 1)).squeeze(-1), y)
        dense.zero_grad(); loss.backward()
        ps.push_gradients(client, [user, item])                              # asynchronous pushes, no barrier between workers
        losses.append(loss.item())
    client.wait()
    client.pull_dense()
    moved = (dense.weight.detach() - w0).abs().max().item()
    # every id this worker used is now admitted somewhere, with frequency >= its own occurrence count
    g = torch.Generator().manual_seed(10 + j)
    mine_u = torch.cat([torch.randint(0, 50, (32,), generator=g) if k % 2 == 0 else torch.randint(0, 80, (32,), generator=g) for k in range(2 * STEPS)][0::2])
    f = client.frequency("user", torch.unique(mine_u))
    own_cnt = torch.bincount(mine_u, minlength=50)[torch.unique(mine_u)]
    assert (f >= own_cnt).all(), (f, own_cnt)
    if j == 0:
        client.save(os.path.join(tmp, "ckpt"), STEPS)
        # FileSliceSend / FileSliceRecv: a 1 MB + 3 B file travels in 64 KB slices to a server and back
        blob = os.urandom((1 << 20) + 3)
        with open(os.path.join(tmp, "blob.bin"), "wb") as fh:
            fh.write(blob)
        assert client.send_file(1, os.path.join(tmp, "blob.bin"), os.path.join(tmp, "remote", "blob.bin"), slice_bytes=64 << 10) == len(blob)
        assert client.recv_file(1, os.path.join(tmp, "remote", "blob.bin"), os.path.join(tmp, "back.bin"), slice_bytes=100_000) == len(blob)
        assert open(os.path.join(tmp, "back.bin"), "rb").read() == blob
    native = sorted(p for p, c in client._conns.items() if c is not None)        # PS reached over the C++ data plane
    with open(os.path.join(tmp, f"worker{j}.json"), "w") as fh:
        json.dump({"losses": losses, "dense_moved": moved, "native_ps": native, "transport": client.transport}, fh)
    client.shutdown()


import pytest  # noqa: E402


@pytest.mark.parametrize("transport", ["native", "rpc"])
def test_async_ps_training(tmp_path, transport, monkeypatch):
    """``native``: sparse pulls / pushes travel over the C++ data plane (csrc/host/ps_server.cc: TCP, one server thread per worker, applies
    straight on the HostEV engine); ``rpc``: everything over torch.distributed.rpc.  Same training, same checks."""
    monkeypatch.setenv("DEEPREC_PS_TRANSPORT", transport)
    _spawn_roles(_proc, NUM_PS + NUM_WORKERS, str(tmp_path))
    stats = [json.load(open(tmp_path / f"ps{i}.json")) for i in range(NUM_PS)]
    workers = [json.load(open(tmp_path / f"worker{j}.json")) for j in range(NUM_WORKERS)]
    # rows are partitioned over the PS processes (key % 1000 % num_ps), both shards non-empty; every push was applied
    assert all(s["user"] > 0 and s["item"] > 0 for s in stats)
    assert sum(s["user"] for s in stats) <= 50 and sum(s["item"] for s in stats) <= 80
    assert sum(s["pushes"] for s in stats) >= NUM_WORKERS * STEPS * 2
    assert all(w["dense_moved"] > 0 and all(l == l for l in w["losses"]) for w in workers)
    assert all(w["transport"] == transport and w["native_ps"] == (list(range(NUM_PS)) if transport == "native" else []) for w in workers), workers
    # the PS-side checkpoint holds exactly the admitted keys of each shard
    from deeprec_b200.checkpoint.saver import BundleReader
    import glob
    n_user = 0
    for i in range(NUM_PS):
        prefix = glob.glob(str(tmp_path / f"ckpt.ps{i}-*.index"))[0][:-6]
        r = BundleReader(prefix)
        keys = r.read(f"user/part_{i}-keys")
        assert ((keys % 1000 % NUM_PS) == i).all()
        n_user += keys.numel()
    # (async training: the other worker may still be pushing new keys after worker 0's save -- the checkpoint is a snapshot, the stats are final)
    assert 0 < n_user <= sum(s["user"] for s in stats)


# ---------------------------------------------------------------------------------------------- elastic PS scaling
E_TOTAL_PS, E_WORKERS = 3, 2


def _elastic_proc(rank, port, tmp):
    torch.manual_seed(0)
    from deeprec_b200.parallel import ps
    if rank < E_TOTAL_PS:
        ps.run_ps(rank, E_TOTAL_PS, E_WORKERS, port, stats_path=os.path.join(tmp, f"eps{rank}.json"), active_ps=2)
        return
    j = rank - E_TOTAL_PS
    client = ps.PSClient(j, E_TOTAL_PS, E_WORKERS, port, active_ps=2)
    emb = client.create_embedding("item", 4, optimizer="adagrad", lr=0.1, seed=5)
    ids = torch.arange(0, 300) * 7 + j            # the two workers touch different keys
    def step():
        rows = ps.group_pull(client, [emb], [ids])[0]
        rows.sum().backward()
        ps.push_gradients(client, [emb])
    for _ in range(2):
        step()
    client.wait()
    flag = os.path.join(tmp, "scaled")
    if j == 0:
        before = client.pull_many([("item", ids)])[0].clone()
        acc_meta = client.fetch_params_meta()
        assert acc_meta[2]["item"][1] == 0 and acc_meta[0]["item"][1] > 0          # the spare server holds nothing yet
        moved = client.scale(3)                                                     # scale UP 2 -> 3
        assert moved > 0 and client.num_ps == 3
        meta = client.fetch_params_meta()
        assert all(m["item"][1] > 0 for m in meta) and sum(m["item"][1] for m in meta) == sum(m["item"][1] for m in acc_meta)
        after = client.pull_many([("item", ids)])[0]
        assert torch.equal(before, after)                                           # rows survive the move bit for bit
        step(); client.wait()                                                       # optimizer slots moved too: training continues
        f = client.frequency("item", ids[:5])
        assert (f == 3).all(), f
        moved_down = client.scale(1)                                                # scale DOWN 3 -> 1
        meta = client.fetch_params_meta()
        assert moved_down > 0 and meta[1]["item"][1] == 0 and meta[2]["item"][1] == 0
        assert (client.frequency("item", ids[:5]) == 3).all()
        open(flag, "w").write("1")
    else:
        # this worker keeps training while the other one re-shards: stale-definition rejections are retried transparently
        import time
        n, t0 = 0, time.time()
        while not os.path.exists(flag) and time.time() - t0 < 150:
            step(); n += 1
            time.sleep(0.01)
        assert os.path.exists(flag), "the scaling worker did not finish"
        client.wait()
        client.refresh_server_def()
        assert client.num_ps == 1
        f = client.frequency("item", ids[:5])
        assert (f == 2 + n).all(), (f, n)                                           # no push was lost or applied twice
    client.shutdown()


def test_elastic_ps_scale_up_and_down(tmp_path):
    _spawn_roles(_elastic_proc, E_TOTAL_PS + E_WORKERS, str(tmp_path), 300)
    stats = [json.load(open(tmp_path / f"eps{i}.json")) for i in range(E_TOTAL_PS)]
    assert stats[0]["item"] == 600 and stats[1]["item"] == 0 and stats[2]["item"] == 0


def test_ps_train_launcher_spawns_roles_and_scales(tmp_path):
    """`python -m deeprec_b200.parallel.ps_train --spawn`: 1 PS (+1 spare) and 2 workers, elastic re-shard onto 2 servers at step 6."""
    import subprocess
    import sys
    res = str(tmp_path / "res")
    r = subprocess.run([sys.executable, "-m", "deeprec_b200.parallel.ps_train", "--spawn", "--num_ps", "1", "--spare_ps", "1", "--num_workers", "2",
                        "--steps", "14", "--batch_size", "128", "--scale_at", "6:2", "--log_every", "0", "--result", res],
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "re-sharded onto 2 servers" in r.stdout
    outs = [json.load(open(f"{res}.worker{j}.json")) for j in range(2)]
    assert all(o["last_loss"] < o["first_loss"] for o in outs) and outs[0]["active_ps"] == 2
