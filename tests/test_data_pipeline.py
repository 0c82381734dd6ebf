"""Staging buffer / prefetch / WorkQueue / Parquet (prefetch_test.py, work_queue_test.py, parquet_dataset tests)."""
import threading
import time

import pytest
import torch

from deeprec_b200.data import PackedHostBatch, ParquetDataset, StagingBuffer, WorkQueue, criteo_batch, smart_stage, staged, taobao_batch


def test_staging_buffer_capacity_blocks_and_close():
    buf = StagingBuffer(capacity=2)
    assert buf.put(1) and buf.put(2) and buf.size() == 2
    with pytest.raises(TimeoutError):
        buf.put(3, timeout_millis=50)                    # full
    assert buf.take() == 1
    t = threading.Thread(target=lambda: (time.sleep(0.05), buf.put(3)))
    t.start()
    assert buf.take() == 2 and buf.take(timeout_millis=2000) == 3
    t.join()
    with pytest.raises(TimeoutError):
        buf.take(timeout_millis=30)                      # empty
    buf.put(4); assert buf.cancel() == 1 and buf.size() == 0
    assert buf.put(5) is False                           # cancelled: producers rejected
    buf.resume(); assert buf.put(6)
    buf.close()
    assert buf.take() == 6
    with pytest.raises(StopIteration):
        buf.take()


def test_staged_iterates_everything_in_order_with_preprocess():
    data = [(torch.full((4, 3), float(i)), torch.arange(4) + i) for i in range(20)]
    st = staged(data, capacity=3, num_threads=1, preprocess=lambda b: (b[0] * 2, b[1]))
    got = list(st)
    assert len(got) == 20
    for i, (x, y) in enumerate(got):
        assert torch.all(x == 2.0 * i) and y[0].item() == i
    st.close()


def test_staged_callable_producer_multithreaded_and_dict_batches():
    counter = {"n": 0}
    lock = threading.Lock()

    def produce():
        with lock:
            if counter["n"] >= 50:
                raise StopIteration
            counter["n"] += 1
            k = counter["n"]
        return {"dense": torch.ones(2, 2) * k, "ids": torch.tensor([k])}
    st = staged(produce, capacity=4, num_threads=3)
    seen = sorted(int(b["ids"].item()) for b in st)
    assert seen == list(range(1, 51))


def test_packed_host_batch_roundtrip():
    ts = [torch.randn(5, 3), torch.arange(7, dtype=torch.int64), torch.tensor([1.5]), torch.empty(0)]
    p = PackedHostBatch(ts, pin=False)
    for a, b in zip(ts, p.unpack_host()):
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    assert p.nbytes % 256 == 0


def test_work_queue_epochs_shuffle_and_resume():
    works = [f"file{i}" for i in range(7)]
    q = WorkQueue(works, num_epochs=2, shuffle=True, seed=3)
    first = [q.take() for _ in range(5)]
    st = q.state_dict()
    rest = list(q.input_producer())
    assert len(first + rest) == 14 and sorted(first + rest) == sorted(works * 2)
    assert sorted((first + rest)[:7]) == sorted(works)              # each epoch is a permutation
    q2 = WorkQueue(works, num_epochs=2, shuffle=True, seed=3)
    q2.load_state_dict(st)
    assert list(q2.input_producer()) == rest                        # resumable position
    q3 = WorkQueue(["a", "b"], num_epochs=1, shuffle=False, num_slices=2)
    assert list(q3.input_producer()) == ["a?slice=0/2", "a?slice=1/2", "b?slice=0/2", "b?slice=1/2"]


def test_parquet_dataset_dense_ragged_and_fields(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq
    n = 100
    tbl = pa.table({"label": pa.array([i % 2 for i in range(n)], pa.int32()), "I1": pa.array([float(i) for i in range(n)], pa.float32()),
                    "C1": pa.array([i * 7 for i in range(n)], pa.int64()), "hist": pa.array([[j for j in range(i % 4)] for i in range(n)], pa.list_(pa.int64())),
                    "tag": pa.array([f"t{i % 5}" for i in range(n)])})
    path = str(tmp_path / "d.parquet")
    pq.write_table(tbl, path, row_group_size=32)
    ds = ParquetDataset(path, batch_size=16, fields=["label", "C1", "hist", "tag"])
    rows = 0
    for b in ds:
        assert set(b) == {"label", "C1", "hist", "tag"}
        rows += b["label"].numel()
        h = b["hist"]
        assert h.nested_row_splits[0][-1].item() == h.values.numel()
        sp = h.to_sparse()
        assert sp.batch_size == b["label"].numel()
        assert b["tag"].dtype == torch.int64
    assert rows == n
    assert sum(b["label"].numel() for b in ParquetDataset(path, batch_size=16, drop_remainder=True)) <= n


def test_synthetic_generators_shapes_and_skew():
    d, ids, y = criteo_batch(4096, 13, [1000, 50, 1000000], seed=1, alpha=1.05)
    assert d.shape == (4096, 13) and ids.shape == (3, 4096) and y.shape == (4096,)
    assert int(ids[0].max()) < 1000 and int(ids[2].max()) < 1000000 and ids.min().item() >= 0
    assert ids[2].unique().numel() < 4096                       # power-law => duplicates
    d2, ids2, _ = criteo_batch(4096, 13, [1000, 50, 1000000], seed=1, alpha=1.05)
    assert torch.equal(ids, ids2) and torch.equal(d, d2)         # deterministic
    tb = taobao_batch(64, max_len=10)
    assert tb["hist_item"].shape == (64, 10) and int(tb["hist_len"].max()) <= 10
    assert torch.all((tb["hist_item"] >= 0).sum(1) == tb["hist_len"])


def test_smart_stage_cpu_passthrough():
    data = [(torch.ones(2) * i,) for i in range(5)]
    out = [b[0][0].item() for b in smart_stage(data, device=None)]
    assert out == [0.0, 1.0, 2.0, 3.0, 4.0]


def _wq_proc(rank, port, tmp):
    import datetime
    import json
    import time
    import torch.distributed as dist
    from deeprec_b200.data import WorkQueue
    store = dist.TCPStore("127.0.0.1", port, 2, is_master=(rank == 0), timeout=datetime.timedelta(seconds=60))
    wq = WorkQueue([f"part-{i:03d}" for i in range(40)], num_epochs=2, shuffle=True, seed=7, store=store, rank=rank, name="wq_test")
    wq.start_service()
    store.set(f"ready{rank}", "1")
    store.wait(["ready0", "ready1"])                       # both consumers exist before the first item is taken (process start-up is not part of the test)
    got = []
    while True:
        w = wq.take()
        if w is None:
            break
        got.append(w)
        time.sleep(0.001 if rank == 0 else 0.004)          # rank 1 is a straggler: it simply ends up with fewer items
    with open(f"{tmp}/wq{rank}.json", "w") as f:
        json.dump(got, f)
    store.set(f"done{rank}", "1")
    store.wait(["done0", "done1"])                         # rank 0 keeps serving until everybody is done
    wq.stop_service()


def test_work_queue_shared_between_processes(tmp_path):
    import json
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_wq_proc, args=(r, port, str(tmp_path))) for r in range(2)]
    [p.start() for p in procs]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    a, b = json.load(open(tmp_path / "wq0.json")), json.load(open(tmp_path / "wq1.json"))
    assert len(a) + len(b) == 80 and len(a) > 0 and len(b) > 0                          # every item of both epochs exactly once, both workers took part
    from collections import Counter
    assert set(Counter(a + b).values()) == {2} and len(set(a + b)) == 40


def test_kafka_dataset_with_an_in_memory_broker():
    """KafkaDataset (docs/docs_en/KafkaDataset.md): subscriptions "topic:partition:offset:length" are drained in order, eof stops at the
    end of a partition, message keys on request, and the saved position resumes exactly after the last delivered message."""
    from deeprec_b200.data import KafkaDataset

    class Broker:
        def __init__(self):
            self.logs = {("clicks", 0): [(b"k%d" % i, b"c0-%d" % i) for i in range(25)], ("clicks", 1): [(None, b"c1-%d" % i) for i in range(7)]}
            self.polls = 0

        def __call__(self, servers, group, config):
            self.servers, self.group, self.config = servers, group, config
            return self

        def poll(self, topic, partition, offset, max_records, timeout_ms):
            self.polls += 1
            log = self.logs[(topic, partition)]
            return [(o, log[o][0], log[o][1]) for o in range(offset, min(len(log), offset + max_records))]

    broker = Broker()
    ds = KafkaDataset(["clicks:0:5:17", "clicks:1"], servers="b1:9092", group="g", eof=True, timeout=10, config_global=["enable.auto.commit=false"],
                      config_topic=["auto.offset.reset=earliest"], consumer_factory=broker, max_poll_records=4)
    assert broker.servers == ["b1:9092"] and broker.group == "g" and broker.config == {"enable.auto.commit": "false", "auto.offset.reset": "earliest"}
    got = list(ds)
    assert got == [b"c0-%d" % i for i in range(5, 17)] + [b"c1-%d" % i for i in range(7)]            # offsets [5, 17) of partition 0 (4th field = exclusive END offset, as in the reference), then all of 1
    assert ds.positions() == [("clicks", 0, 17), ("clicks", 1, 7)]
    # batches + keys + resume from a saved position
    ds2 = KafkaDataset(["clicks:0:0:-1"], eof=True, message_key=True, consumer_factory=Broker(), max_poll_records=10)
    it = ds2.batch(8, parse_fn=lambda ms: [v.decode() for _, v in ms])
    first = next(it)
    assert first == [f"c0-{i}" for i in range(8)]
    state = ds2.state_dict()
    ds3 = KafkaDataset(["clicks:0:0:-1"], eof=True, message_key=True, consumer_factory=Broker())
    ds3.load_state_dict(state)
    rest = [m for b in ds3.batch(8) for m in b]
    assert [k for k, _ in rest] == [b"k%d" % i for i in range(8, 25)] and len(rest) == 17
    # a position saved while a batch is only partially filled is the last BATCH boundary: nothing is dropped on restore
    ds4 = KafkaDataset(["clicks:0:0:20"], eof=True, consumer_factory=Broker(), max_poll_records=3)
    it4 = ds4.batch(8)
    next(it4)
    st4 = ds4.state_dict()
    assert st4["subscriptions"] == ["clicks:0:8:20"]
    ds5 = KafkaDataset(["clicks:0:0:20"], eof=True, consumer_factory=Broker())
    ds5.load_state_dict(st4)
    assert [m for b in ds5.batch(8) for m in b] == [b"c0-%d" % i for i in range(8, 20)]
    # without a client library and without a factory the built-in wire-protocol consumer is used (tests/test_kafka_wire.py drives it against a TCP broker)
    try:
        import kafka  # noqa: F401
    except ImportError:
        assert type(KafkaDataset(["t"])._consumer).__name__ == "KafkaWireConsumer"
