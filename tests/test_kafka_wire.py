"""data/kafka_wire.py: the built-in Kafka wire-protocol consumer against an in-process TCP broker (tests/kafka_mini_broker.py), and
KafkaDataset on top of it (reference: contrib/kafka/kernels/kafka_dataset_ops.cc -- librdkafka consumer inside the dataset kernel)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kafka_mini_broker import MiniKafkaBroker  # noqa: E402

from deeprec_b200.data.kafka_dataset import KafkaDataset  # noqa: E402
from deeprec_b200.data.kafka_wire import (KafkaProtocolError, KafkaWireConsumer, crc32c, decode_record_set, enc_varint, encode_batch_v2,  # noqa: E402
                                          _Reader)


def test_crc32c_and_varints():
    assert crc32c(b"123456789") == 0xE3069283                        # the standard CRC-32C check value
    for v in (0, 1, -1, 63, -64, 64, 300, -301, 2 ** 31, -(2 ** 40)):
        assert _Reader(enc_varint(v)).varint() == v


@pytest.mark.parametrize("compress", [False, True])
def test_batch_round_trip_and_corruption(compress):
    recs = [(None, b"a"), (b"k", b""), (b"key2", b"x" * 300)]
    raw = encode_batch_v2(41, recs, compress)
    assert decode_record_set(raw) == [(41, None, b"a"), (42, b"k", b""), (43, b"key2", b"x" * 300)]
    assert decode_record_set(raw + raw[:30]) == decode_record_set(raw)           # truncated trailing batch is ignored
    bad = bytearray(raw); bad[-1] ^= 1
    with pytest.raises(KafkaProtocolError):
        decode_record_set(bytes(bad))
    if compress:                                                                 # the gzip trailer catches it even without the batch CRC
        with pytest.raises(KafkaProtocolError):
            decode_record_set(bytes(bad), check_crc=False)
    else:
        assert len(decode_record_set(bytes(bad), check_crc=False)) == 3


@pytest.mark.parametrize("compress,truncate", [(False, False), (True, True)])
def test_consumer_against_the_tcp_broker(compress, truncate):
    b = MiniKafkaBroker(compress=compress, truncate_last=truncate)
    try:
        for i in range(0, 30, 3):
            b.append("clicks", 0, [(f"k{j}".encode(), f"v{j}".encode()) for j in range(i, i + 3)])
        b.append("clicks", 1, [(None, b"p1")])
        c = KafkaWireConsumer(f"127.0.0.1:{b.port}", "", {"client.id": "t"})
        got = c.poll("clicks", 0, 4, 100, 500)                                     # offset 4 sits inside the second batch
        assert [m[0] for m in got] == list(range(4, 30)) and got[0] == (4, b"k4", b"v4")
        assert [m[0] for m in c.poll("clicks", 0, 10, 5, 500)] == [10, 11, 12, 13, 14]      # max_records
        assert c.poll("clicks", 1, 0, 10, 200) == [(0, None, b"p1")]
        assert c.poll("clicks", 0, 30, 10, 100) == []                              # at the log end: long poll times out empty
        assert c.list_offset("clicks", 0, -2) == 0 and c.list_offset("clicks", 0, -1) == 30
        assert [m[0] for m in c.poll("clicks", 0, -2, 2, 200)] == [0, 1]            # EARLIEST
        with pytest.raises(KafkaProtocolError):
            c.poll("nope", 0, 0, 1, 100)
        b.fail_next_fetch_with = 6                                                 # NOT_LEADER once: metadata refresh + retry
        assert [m[0] for m in c.poll("clicks", 0, 28, 10, 300)] == [28, 29]
        assert [m[0] for m in c.poll("clicks", 0, 99, 3, 300)] == [0, 1, 2]         # out of range -> auto.offset.reset = earliest
        c.close()
    finally:
        b.close()


def test_kafka_dataset_over_the_wire_with_saved_position():
    b = MiniKafkaBroker()
    try:
        for i in range(0, 20, 4):
            b.append("train", 0, [(None, f"{j},{j * 2}".encode()) for j in range(i, i + 4)])
        b.append("train", 1, [(None, b"100,200"), (None, b"101,202")])
        servers = f"127.0.0.1:{b.port}"
        ds = KafkaDataset(["train:0:2:17", "train:1"], servers=servers, eof=True, timeout=200)     # default client = the wire consumer (no kafka-python here)
        assert type(ds._consumer).__name__ in ("KafkaWireConsumer", "_KafkaPythonConsumer")
        it = ds.batch(5, parse_fn=lambda ms: [int(m.split(b",")[0]) for m in ms])
        first = next(it)
        assert first == [2, 3, 4, 5, 6]
        state = ds.state_dict()
        rest = [x for bt in it for x in bt]
        assert rest == list(range(7, 17)) + [100, 101]
        ds2 = KafkaDataset(["train:0:2:17", "train:1"], servers=servers, eof=True, timeout=200)
        ds2.load_state_dict(state)
        assert [int(m.split(b",")[0]) for m in ds2] == list(range(7, 17)) + [100, 101]
    finally:
        b.close()


def test_group_dataset_splits_partitions_and_survives_a_resize():
    """Two workers split four partitions; after a resize to three workers every message is still delivered exactly once."""
    from deeprec_b200.data.kafka_dataset import KafkaGroupIODataset, merge_group_states
    b = MiniKafkaBroker()
    try:
        for p in range(4):
            b.append("ev", p, [(None, f"{p}-{i}".encode()) for i in range(6)])
        servers = f"127.0.0.1:{b.port}"
        ws = [KafkaGroupIODataset(["ev"], servers=servers, world_size=2, rank=r, eof=True, timeout=100, max_poll_records=2) for r in range(2)]
        assert [sorted(p for _, p in w.rebalance(2, r)) for r, w in enumerate(ws)] == [[0, 2], [1, 3]]
        seen = []
        for w in ws:                                                   # every worker consumes 5 messages, then the job is resized
            it = iter(w)
            seen += [next(it) for _ in range(5)]
        merged = merge_group_states([w.state_dict() for w in ws])
        ws3 = [KafkaGroupIODataset(["ev"], servers=servers, world_size=3, rank=r, eof=True, timeout=100) for r in range(3)]
        for w in ws3:
            w.load_state_dict(merged)
            seen += list(w)
        assert sorted(seen) == sorted(f"{p}-{i}".encode() for p in range(4) for i in range(6))
        with pytest.raises(ValueError):
            ws[0].rebalance(2, 2)
    finally:
        b.close()
