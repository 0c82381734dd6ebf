"""Streaming input for online learning (reference: docs/docs_en/KafkaDataset.md -- ``KafkaDataset(topics, servers, group, eof, timeout,
config_global, config_topic, message_key)`` with saveable position).

Subscriptions use the reference's ``"topic:partition:offset:length"`` strings where, as in the reference kernel
(contrib/kafka/kernels/kafka_dataset_ops.cc:121: reading stops when ``offset >= limit``), the 4th field is an ABSOLUTE, EXCLUSIVE end offset
(-1 = unlimited): ``"t:0:100:200"`` delivers offsets 100..199.  Messages are consumed partition by partition in offset order, the position of
every subscription is part of ``state_dict()`` so a restored job continues exactly after the last message it delivered -- when the stream is
consumed through ``batch(n, parse_fn)`` the saved position is the one at the last BATCH boundary, so messages sitting in a partially filled
batch are read again after a restore instead of being dropped.

The broker client is pluggable: ``kafka-python`` (``kafka.KafkaConsumer``) is used when it is installed, otherwise the built-in wire-protocol
consumer (:mod:`data.kafka_wire`: Metadata / ListOffsets / Fetch over TCP, record-batch v2 + legacy message sets, CRC-32C, gzip) -- and
``consumer_factory`` accepts anything with the small interface below (the tests drive the dataset with an in-memory broker AND with an
in-process TCP broker speaking the wire protocol):

    consumer = consumer_factory(servers, group, config)      # config: dict from config_global / config_topic "key=value" strings
    consumer.poll(topic, partition, offset, max_records, timeout_ms) -> list[(offset, key: bytes | None, value: bytes)]
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple


@dataclass
class _Subscription:
    topic: str
    partition: int
    offset: int           # next offset to read
    limit: int            # exclusive end offset, -1 = unlimited


def _parse_subscription(s: str) -> _Subscription:
    parts = s.split(":")
    if not parts[0]:
        raise ValueError(f"bad subscription {s!r}: expected topic[:partition[:offset[:length]]]")
    nums = [int(p) for p in parts[1:4]] + [0, 0, -1][len(parts) - 1:]
    return _Subscription(parts[0], nums[0], nums[1], nums[2])


class _KafkaPythonConsumer:
    """Adapter over kafka-python (only imported when no consumer_factory is given)."""

    def __init__(self, servers, group, config):
        try:
            from kafka import KafkaConsumer, TopicPartition  # type: ignore
        except ImportError as e:  # pragma: no cover - no client library in this image
            raise ImportError("KafkaDataset needs a Kafka client: install kafka-python or pass consumer_factory=...") from e
        self._tp = TopicPartition
        kw = {k.replace(".", "_"): v for k, v in config.items()}
        kw.setdefault("enable_auto_commit", False)
        self._c = KafkaConsumer(bootstrap_servers=servers, group_id=group or None, **kw)
        self._assigned = None

    def poll(self, topic, partition, offset, max_records, timeout_ms):  # pragma: no cover - needs a broker
        tp = self._tp(topic, partition)
        if self._assigned != tp:
            self._c.assign([tp]); self._assigned = tp
        self._c.seek(tp, offset)
        got = self._c.poll(timeout_ms=timeout_ms, max_records=max_records).get(tp, [])
        return [(m.offset, m.key, m.value) for m in got]


def _default_consumer():
    try:
        import kafka  # type: ignore  # noqa: F401
        return _KafkaPythonConsumer
    except ImportError:
        from .kafka_wire import KafkaWireConsumer
        return KafkaWireConsumer


class KafkaDataset:
    def __init__(self, topics: Sequence[str], servers="localhost", group: str = "", eof: bool = False, timeout: int = 1000,
                 config_global: Optional[Sequence[str]] = None, config_topic: Optional[Sequence[str]] = None, message_key: bool = False,
                 consumer_factory: Optional[Callable] = None, max_poll_records: int = 500):
        self.subs: List[_Subscription] = [_parse_subscription(t) for t in ([topics] if isinstance(topics, str) else topics)]
        self.eof, self.timeout, self.message_key, self.max_poll_records = bool(eof), int(timeout), bool(message_key), int(max_poll_records)
        config: Dict[str, str] = {}
        for kv in list(config_global or []) + list(config_topic or []):
            k, _, v = kv.partition("=")
            config[k.strip()] = v.strip()
        servers = [servers] if isinstance(servers, str) else list(servers)
        self._consumer = (consumer_factory or _default_consumer())(servers, group, config)
        self._cur = 0                                   # subscription being drained (the reference reads them in order)

    # ---- iteration ------------------------------------------------------------------------------------------------------------
    def __iter__(self) -> Iterator:
        while self._cur < len(self.subs):
            s = self.subs[self._cur]
            if s.limit >= 0 and s.offset >= s.limit:
                self._cur += 1
                continue
            want = self.max_poll_records if s.limit < 0 else min(self.max_poll_records, s.limit - s.offset)
            msgs = self._consumer.poll(s.topic, s.partition, s.offset, want, self.timeout)
            if not msgs:
                if self.eof:                            # end of this partition: move on; otherwise keep waiting for new messages
                    self._cur += 1
                continue
            for off, key, value in msgs:
                if s.limit >= 0 and off >= s.limit:
                    s.offset = s.limit
                    break
                s.offset = off + 1                      # position advances BEFORE the message is handed out: a checkpoint taken by the
                yield (key, value) if self.message_key else value      # consumer of this generator never replays what it already received

    def batch(self, batch_size: int, parse_fn: Optional[Callable] = None, drop_remainder: bool = False) -> Iterator:
        buf = []
        self._committed = self._snapshot()
        for m in self:
            buf.append(m)
            if len(buf) == batch_size:
                self._committed = self._snapshot()      # batch boundary: everything up to here has been handed to the trainer
                yield parse_fn(buf) if parse_fn else buf
                buf = []
        if buf and not drop_remainder:
            self._committed = self._snapshot()
            yield parse_fn(buf) if parse_fn else buf
        self._committed = None

    # ---- saveable position (make_saveable_from_iterator in the reference) ---------------------------------------------------------
    def _snapshot(self) -> dict:
        return {"current": self._cur, "subscriptions": [f"{s.topic}:{s.partition}:{s.offset}:{s.limit}" for s in self.subs]}

    def state_dict(self) -> dict:
        committed = getattr(self, "_committed", None)
        return dict(committed) if committed is not None else self._snapshot()

    def load_state_dict(self, state: dict) -> None:
        self.subs = [_parse_subscription(t) for t in state["subscriptions"]]
        self._cur = int(state["current"])

    def positions(self) -> List[Tuple[str, int, int]]:
        return [(s.topic, s.partition, s.offset) for s in self.subs]
