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


class KafkaGroupIODataset(KafkaDataset):
    """Group consumption for data-parallel online learning (reference: ``KafkaGroupIODataset`` -- the consumers of one group split the
    partitions of the subscribed topics among themselves and re-split when the group changes).

    Here the group is the training job: worker ``rank`` of ``world_size`` reads the partitions ``p`` with ``p % world_size == rank`` of every
    topic, discovered from the broker's metadata (``consumer.partitions(topic)``) -- the assignment a range / round-robin assignor converges
    to for a static membership, without a coordinator round trip.  ``rebalance(world_size, rank)`` re-splits after an elastic resize;
    positions travel with the partitions through ``state_dict()`` (merge the workers' states with :func:`merge_group_states` before the resize,
    load the merged state on every new worker: each keeps the positions of the partitions it now owns)."""

    def __init__(self, topics: Sequence[str], servers="localhost", group: str = "", world_size: int = 1, rank: int = 0, offset: int = 0, **kw):
        self._topics = [topics] if isinstance(topics, str) else list(topics)
        self._start = int(offset)
        super().__init__([], servers=servers, group=group, **kw)
        self._known: Dict[Tuple[str, int], int] = {}          # (topic, partition) -> next offset, for every partition this worker ever saw or was told about
        self.rebalance(world_size, rank)

    def _all_partitions(self) -> List[Tuple[str, int]]:
        if not hasattr(self._consumer, "partitions"):
            raise TypeError("KafkaGroupIODataset needs a broker client with partitions(topic) (data.kafka_wire.KafkaWireConsumer has it)")
        return [(t, p) for t in self._topics for p in self._consumer.partitions(t)]

    def rebalance(self, world_size: int, rank: int) -> List[Tuple[str, int]]:
        """Re-split the partitions for a group of ``world_size`` workers; returns this worker's assignment."""
        if not 0 <= rank < world_size:
            raise ValueError(f"rank {rank} outside a group of {world_size}")
        for s in self.subs:
            self._known[(s.topic, s.partition)] = s.offset
        self.world_size, self.rank = int(world_size), int(rank)
        mine = [(t, p) for i, (t, p) in enumerate(sorted(self._all_partitions())) if i % world_size == rank]
        self.subs = [_Subscription(t, p, self._known.get((t, p), self._start), -1) for t, p in mine]
        self._cur = 0
        return mine

    def state_dict(self) -> dict:
        st = super().state_dict()
        known = dict(self._known)
        for sub in st["subscriptions"]:
            s = _parse_subscription(sub)
            known[(s.topic, s.partition)] = s.offset
        st["group_positions"] = {f"{t}:{p}": o for (t, p), o in sorted(known.items())}
        return st

    def load_state_dict(self, state: dict) -> None:
        for k, o in state.get("group_positions", {}).items():
            t, _, p = k.rpartition(":")
            self._known[(t, int(p))] = int(o)
        for s in self.subs:
            s.offset = self._known.get((s.topic, s.partition), s.offset)
        self._cur = 0


def merge_group_states(states: Sequence[dict]) -> dict:
    """Union of the per-worker positions of one consumer group (every partition is owned by exactly one worker at a time: the furthest
    position wins where a partition changed hands)."""
    pos: Dict[str, int] = {}
    for st in states:
        for k, o in st.get("group_positions", {}).items():
            pos[k] = max(int(o), pos.get(k, -1))
    return {"current": 0, "subscriptions": [], "group_positions": pos}
