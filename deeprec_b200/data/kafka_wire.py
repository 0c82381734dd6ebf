"""A Kafka consumer speaking the broker wire protocol directly over TCP (no client library): Metadata (leader discovery), ListOffsets
(earliest / latest) and Fetch, with record-batch v2 (magic 2, varint records, CRC-32C, gzip) and legacy message-set (magic 0 / 1) decoding.

It is the default broker client of :class:`data.kafka_dataset.KafkaDataset` when ``kafka-python`` is not installed -- the reference links
librdkafka into its dataset kernel (``contrib/kafka/kernels/kafka_dataset_ops.cc``: ``RdKafka::KafkaConsumer::consume`` per message); a
training input reader needs only the read path of the protocol, which is small enough to own.

Interface (what KafkaDataset calls):  ``poll(topic, partition, offset, max_records, timeout_ms) -> [(offset, key | None, value)]``.
Not implemented: consumer groups / offset commits (the dataset's position is saved in the training checkpoint, as in the reference),
SASL / TLS, snappy / lz4 / zstd batches (a clear error names the codec)."""
from __future__ import annotations

import gzip
import socket
import struct
import time
import zlib
from typing import Dict, List, Optional, Tuple

API_FETCH, API_LIST_OFFSETS, API_METADATA = 1, 2, 3
EARLIEST, LATEST = -2, -1
_ERRORS = {1: "OFFSET_OUT_OF_RANGE", 3: "UNKNOWN_TOPIC_OR_PARTITION", 5: "LEADER_NOT_AVAILABLE", 6: "NOT_LEADER_FOR_PARTITION", 9: "REPLICA_NOT_AVAILABLE"}


class KafkaProtocolError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
def _crc32c_table():
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t.append(c)
    return t


_CRC_T = _crc32c_table()


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_T[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------------- primitive codecs
class _Reader:
    def __init__(self, buf: bytes, pos: int = 0, end: Optional[int] = None):
        self.b, self.p, self.e = buf, pos, len(buf) if end is None else end

    def left(self) -> int:
        return self.e - self.p

    def take(self, n: int) -> bytes:
        if n < 0 or self.p + n > self.e:
            raise EOFError
        v = self.b[self.p:self.p + n]; self.p += n
        return v

    def i8(self): return struct.unpack(">b", self.take(1))[0]
    def i16(self): return struct.unpack(">h", self.take(2))[0]
    def i32(self): return struct.unpack(">i", self.take(4))[0]
    def u32(self): return struct.unpack(">I", self.take(4))[0]
    def i64(self): return struct.unpack(">q", self.take(8))[0]

    def string(self) -> Optional[str]:
        n = self.i16()
        return None if n < 0 else self.take(n).decode()

    def bytes_(self) -> Optional[bytes]:
        n = self.i32()
        return None if n < 0 else self.take(n)

    def varint(self) -> int:                       # zig-zag varint (records of a v2 batch)
        shift = result = 0
        while True:
            b = self.take(1)[0]
            result |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
            if shift > 63:
                raise KafkaProtocolError("varint too long")
        return (result >> 1) ^ -(result & 1)

    def vbytes(self) -> Optional[bytes]:
        n = self.varint()
        return None if n < 0 else self.take(n)


def enc_string(s: Optional[str]) -> bytes:
    if s is None:
        return struct.pack(">h", -1)
    b = s.encode()
    return struct.pack(">h", len(b)) + b


def enc_varint(v: int) -> bytes:
    v = (v << 1) ^ (v >> 63)
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


# ---------------------------------------------------------------------------------------------------------------- record decoding
Message = Tuple[int, Optional[bytes], bytes]


def decode_record_set(buf: bytes, check_crc: bool = True) -> List[Message]:
    """All complete messages of a fetch response's record set, in offset order.  A truncated trailing batch (the broker cuts at
    ``max_bytes``) is ignored."""
    out: List[Message] = []
    r = _Reader(buf)
    while r.left() >= 17:                                            # offset (8) + length (4) + (epoch | crc) (4) + magic (1)
        start = r.p
        base, length = r.i64(), r.i32()
        if length <= 0 or r.left() < length:
            break                                                    # partial batch at the end of the buffer
        end = r.p + length
        magic = buf[start + 16]
        try:
            if magic == 2:
                out.extend(_decode_batch_v2(buf, r.p, end, base, check_crc))
            elif magic in (0, 1):
                out.extend(_decode_legacy(buf, start, end, check_crc))
            else:
                raise KafkaProtocolError(f"unknown record magic {magic}")
        except (EOFError, OSError, zlib.error) as e:                  # truncated fields, bad gzip stream
            raise KafkaProtocolError(f"corrupt record batch at offset {base}: {e}") from e
        r.p = end
    return out


def _decode_batch_v2(buf: bytes, pos: int, end: int, base: int, check_crc: bool) -> List[Message]:
    r = _Reader(buf, pos, end)
    r.i32()                                                          # partition leader epoch
    r.i8()                                                           # magic
    crc = r.u32()
    if check_crc and crc32c(buf[r.p:end]) != crc:
        raise KafkaProtocolError(f"CRC-32C mismatch in the batch at offset {base}")
    attrs = r.i16()
    r.i32(); r.i64(); r.i64(); r.i64(); r.i16(); r.i32()             # lastOffsetDelta, first / max timestamp, producer id / epoch, base sequence
    n = r.i32()
    if attrs & 0x20:                                                 # control batch (transaction markers): no user records
        return []
    codec = attrs & 0x7
    body = buf[r.p:end]
    if codec == 1:
        body = gzip.decompress(body)
    elif codec != 0:
        raise KafkaProtocolError(f"compression codec {['none', 'gzip', 'snappy', 'lz4', 'zstd'][codec] if codec < 5 else codec} is not supported by the built-in client")
    rr = _Reader(body)
    out = []
    for _ in range(n):
        ln = rr.varint()
        rec = _Reader(body, rr.p, rr.p + ln)
        rr.take(ln)
        rec.i8(); rec.varint()                                       # attributes, timestamp delta
        off = base + rec.varint()
        key = rec.vbytes(); val = rec.vbytes()
        out.append((off, key, val if val is not None else b""))      # headers are ignored
    return out


def _decode_legacy(buf: bytes, start: int, end: int, check_crc: bool) -> List[Message]:
    r = _Reader(buf, start, end)
    off, _size = r.i64(), r.i32()
    crc = r.u32()
    if check_crc:
        if zlib.crc32(buf[r.p:end]) & 0xFFFFFFFF != crc:
            raise KafkaProtocolError(f"CRC mismatch in the message at offset {off}")
    magic, attrs = r.i8(), r.i8()
    if magic == 1:
        r.i64()
    key, val = r.bytes_(), r.bytes_()
    codec = attrs & 0x7
    if codec == 0:
        return [(off, key, val if val is not None else b"")]
    if codec != 1:
        raise KafkaProtocolError("legacy message sets: only gzip wrappers are supported")
    inner = decode_record_set(gzip.decompress(val or b""), check_crc)
    if magic == 1 and inner:                                         # magic 1 wrappers carry the LAST offset, inner offsets are relative
        shift = off - inner[-1][0]
        inner = [(o + shift, k, v) for o, k, v in inner]
    return inner


def encode_batch_v2(base_offset: int, records: List[Tuple[Optional[bytes], bytes]], compress: bool = False, timestamp_ms: int = 0) -> bytes:
    """One record batch (producer side of the format; used by the in-process test broker and by tools that write replay files)."""
    body = bytearray()
    for i, (k, v) in enumerate(records):
        rec = bytearray([0]) + enc_varint(0) + enc_varint(i)
        rec += enc_varint(-1) if k is None else enc_varint(len(k)) + k
        rec += enc_varint(len(v)) + v + enc_varint(0)
        body += enc_varint(len(rec)) + rec
    payload = gzip.compress(bytes(body)) if compress else bytes(body)
    after_crc = struct.pack(">hiqqqhii", 1 if compress else 0, len(records) - 1, timestamp_ms, timestamp_ms, -1, -1, -1, len(records)) + payload
    head = struct.pack(">ibI", 0, 2, crc32c(after_crc))
    return struct.pack(">qi", base_offset, len(head) + len(after_crc)) + head + after_crc


# ---------------------------------------------------------------------------------------------------------------- the consumer
class KafkaWireConsumer:
    def __init__(self, servers, group: str = "", config: Optional[Dict[str, str]] = None):
        self.servers = [self._hostport(s) for s in ([servers] if isinstance(servers, str) else servers)]
        cfg = dict(config or {})
        self.client_id = cfg.get("client.id", "deeprec_b200")
        self.max_bytes = int(cfg.get("fetch.max.bytes", cfg.get("max.partition.fetch.bytes", 4 << 20)))
        self.min_bytes = int(cfg.get("fetch.min.bytes", 1))
        self.check_crc = cfg.get("check.crcs", "true").lower() != "false"
        self.reset = cfg.get("auto.offset.reset", "earliest")
        self.sock_timeout = float(cfg.get("socket.timeout.ms", 30000)) / 1e3
        self._conns: Dict[Tuple[str, int], socket.socket] = {}
        self._leaders: Dict[Tuple[str, int], Tuple[str, int]] = {}
        self._corr = 0

    @staticmethod
    def _hostport(s: str) -> Tuple[str, int]:
        host, _, port = s.rpartition(":") if ":" in s else (s, "", "9092")
        return host or "localhost", int(port)

    # ---- transport -----------------------------------------------------------------------------------------------------------------
    def _conn(self, addr) -> socket.socket:
        c = self._conns.get(addr)
        if c is None:
            c = socket.create_connection(addr, timeout=self.sock_timeout)
            c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self._conns[addr] = c
        return c

    def _call(self, addr, api_key: int, version: int, body: bytes, timeout_s: Optional[float] = None) -> _Reader:
        self._corr += 1
        msg = struct.pack(">hhi", api_key, version, self._corr) + enc_string(self.client_id) + body
        try:
            c = self._conn(addr)
            c.settimeout(self.sock_timeout if timeout_s is None else timeout_s + self.sock_timeout)
            c.sendall(struct.pack(">i", len(msg)) + msg)
            size = struct.unpack(">i", self._recv(c, 4))[0]
            data = self._recv(c, size)
        except OSError:
            self._drop(addr)
            raise
        r = _Reader(data)
        if r.i32() != self._corr:
            self._drop(addr)
            raise KafkaProtocolError("correlation id mismatch")
        return r

    @staticmethod
    def _recv(c: socket.socket, n: int) -> bytes:
        buf = bytearray()
        while len(buf) < n:
            chunk = c.recv(n - len(buf))
            if not chunk:
                raise ConnectionError("broker closed the connection")
            buf += chunk
        return bytes(buf)

    def _drop(self, addr) -> None:
        c = self._conns.pop(addr, None)
        if c is not None:
            try:
                c.close()
            except OSError:
                pass

    def close(self) -> None:
        for a in list(self._conns):
            self._drop(a)

    # ---- metadata --------------------------------------------------------------------------------------------------------------------
    def _leader(self, topic: str, partition: int, refresh: bool = False) -> Tuple[str, int]:
        key = (topic, partition)
        if not refresh and key in self._leaders:
            return self._leaders[key]
        last: Optional[Exception] = None
        for addr in self.servers:
            try:
                r = self._call(addr, API_METADATA, 1, struct.pack(">i", 1) + enc_string(topic))
            except OSError as e:
                last = e
                continue
            brokers = {}
            for _ in range(r.i32()):
                node, host, port = r.i32(), r.string(), r.i32()
                r.string()                                               # rack
                brokers[node] = (host, port)
            r.i32()                                                      # controller id
            for _ in range(r.i32()):
                terr, tname = r.i16(), r.string()
                r.i8()                                                   # is_internal
                for _ in range(r.i32()):
                    perr, pid, leader = r.i16(), r.i32(), r.i32()
                    for _ in range(r.i32()):
                        r.i32()                                          # replicas
                    for _ in range(r.i32()):
                        r.i32()                                          # isr
                    if tname == topic and pid == partition:
                        if terr or perr or leader not in brokers:
                            raise KafkaProtocolError(f"{topic}:{partition}: {_ERRORS.get(terr or perr, terr or perr or 'no leader')}")
                        self._leaders[key] = brokers[leader]
            if key in self._leaders:
                return self._leaders[key]
            raise KafkaProtocolError(f"{topic}:{partition}: UNKNOWN_TOPIC_OR_PARTITION")
        raise ConnectionError(f"no bootstrap server reachable: {self.servers}") from last

    def partitions(self, topic: str) -> List[int]:
        """Partition ids of ``topic`` (Metadata request to the first reachable bootstrap server)."""
        last: Optional[Exception] = None
        for addr in self.servers:
            try:
                r = self._call(addr, API_METADATA, 1, struct.pack(">i", 1) + enc_string(topic))
            except OSError as e:
                last = e
                continue
            for _ in range(r.i32()):
                r.i32(); r.string(); r.i32(); r.string()
            r.i32()
            out: List[int] = []
            for _ in range(r.i32()):
                terr, tname = r.i16(), r.string()
                r.i8()
                for _ in range(r.i32()):
                    r.i16(); pid = r.i32(); r.i32()
                    for _ in range(r.i32()):
                        r.i32()
                    for _ in range(r.i32()):
                        r.i32()
                    if tname == topic:
                        out.append(pid)
                if tname == topic and terr:
                    raise KafkaProtocolError(f"{topic}: {_ERRORS.get(terr, terr)}")
            return sorted(out)
        raise ConnectionError(f"no bootstrap server reachable: {self.servers}") from last

    def list_offset(self, topic: str, partition: int, which: int = EARLIEST) -> int:
        """``which``: EARLIEST (-2), LATEST (-1) or a timestamp in ms."""
        body = struct.pack(">ii", -1, 1) + enc_string(topic) + struct.pack(">iiq", 1, partition, which)
        r = self._call(self._leader(topic, partition), API_LIST_OFFSETS, 1, body)
        for _ in range(r.i32()):
            r.string()
            for _ in range(r.i32()):
                pid, err, _ts, off = r.i32(), r.i16(), r.i64(), r.i64()
                if pid == partition:
                    if err:
                        raise KafkaProtocolError(f"ListOffsets {topic}:{partition}: {_ERRORS.get(err, err)}")
                    return off
        raise KafkaProtocolError("ListOffsets: partition missing from the response")

    # ---- fetch -----------------------------------------------------------------------------------------------------------------------
    def poll(self, topic: str, partition: int, offset: int, max_records: int, timeout_ms: int) -> List[Message]:
        deadline = time.monotonic() + timeout_ms / 1e3
        if offset < 0:
            offset = self.list_offset(topic, partition, LATEST if offset == LATEST else EARLIEST)
        retried = False
        while True:
            wait = max(0, int((deadline - time.monotonic()) * 1e3))
            body = (struct.pack(">iiiib", -1, wait, self.min_bytes, self.max_bytes, 0) + struct.pack(">i", 1) + enc_string(topic) +
                    struct.pack(">iiqi", 1, partition, offset, self.max_bytes))
            try:
                r = self._call(self._leader(topic, partition), API_FETCH, 4, body, timeout_s=wait / 1e3)
            except OSError:
                if retried:
                    raise
                retried = True
                self._leader(topic, partition, refresh=True)
                continue
            r.i32()                                                      # throttle time
            records = b""
            err = 0
            for _ in range(r.i32()):
                r.string()
                for _ in range(r.i32()):
                    pid, perr = r.i32(), r.i16()
                    r.i64(); r.i64()                                     # high watermark, last stable offset
                    na = r.i32()
                    for _ in range(max(na, 0)):
                        r.i64(); r.i64()                                 # aborted transactions
                    rs = r.bytes_() or b""
                    if pid == partition:
                        records, err = rs, perr
            if err == 6 and not retried:                                 # NOT_LEADER: refresh the metadata once
                retried = True
                self._leader(topic, partition, refresh=True)
                continue
            if err == 1:                                                 # OFFSET_OUT_OF_RANGE: honour auto.offset.reset
                if self.reset == "none" or retried:
                    raise KafkaProtocolError(f"{topic}:{partition}: offset {offset} out of range")
                retried = True
                offset = self.list_offset(topic, partition, LATEST if self.reset == "latest" else EARLIEST)
                continue
            if err:
                raise KafkaProtocolError(f"Fetch {topic}:{partition}: {_ERRORS.get(err, err)}")
            msgs = [m for m in decode_record_set(records, self.check_crc) if m[0] >= offset]   # batches come whole: skip what precedes the position
            if msgs or time.monotonic() >= deadline:
                return msgs[:max_records]
