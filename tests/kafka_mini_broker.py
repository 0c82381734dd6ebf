"""An in-process TCP broker speaking the subset of the Kafka wire protocol a consumer needs (Metadata v1, ListOffsets v1, Fetch v4 with
record-batch v2), for tests of data/kafka_wire.py.  Logs are appended by the test (``append``); every append call becomes ONE record batch, so
fetching from the middle of a batch returns the whole batch, as a real broker does."""
import socket
import struct
import threading
import time

from deeprec_b200.data.kafka_wire import _Reader, enc_string, encode_batch_v2


class MiniKafkaBroker:
    def __init__(self, compress=False, truncate_last=False):
        self.logs = {}                      # (topic, partition) -> list of (base_offset, n_records, batch_bytes)
        self.next = {}
        self.compress, self.truncate_last = compress, truncate_last
        self.lock = threading.Lock()
        self.srv = socket.socket(); self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind(("127.0.0.1", 0)); self.srv.listen(8)
        self.port = self.srv.getsockname()[1]
        self.stop = False
        self.requests = []
        self.fail_next_fetch_with = 0       # error code injected into the next Fetch response
        threading.Thread(target=self._accept, daemon=True).start()

    def append(self, topic, partition, records):
        with self.lock:
            base = self.next.get((topic, partition), 0)
            self.logs.setdefault((topic, partition), []).append((base, len(records), encode_batch_v2(base, records, self.compress)))
            self.next[(topic, partition)] = base + len(records)

    def close(self):
        self.stop = True
        try:
            self.srv.close()
        except OSError:
            pass

    def _accept(self):
        while not self.stop:
            try:
                c, _ = self.srv.accept()
            except OSError:
                return
            threading.Thread(target=self._serve, args=(c,), daemon=True).start()

    def _serve(self, c):
        try:
            while not self.stop:
                head = self._recv(c, 4)
                if head is None:
                    return
                data = self._recv(c, struct.unpack(">i", head)[0])
                r = _Reader(data)
                api, ver, corr = r.i16(), r.i16(), r.i32()
                r.string()
                self.requests.append((api, ver))
                body = {3: self._metadata, 2: self._list_offsets, 1: self._fetch}[api](r)
                msg = struct.pack(">i", corr) + body
                c.sendall(struct.pack(">i", len(msg)) + msg)
        except (OSError, EOFError):
            pass
        finally:
            c.close()

    @staticmethod
    def _recv(c, n):
        buf = bytearray()
        while len(buf) < n:
            chunk = c.recv(n - len(buf))
            if not chunk:
                return None
            buf += chunk
        return bytes(buf)

    def _metadata(self, r):
        topics = [r.string() for _ in range(r.i32())]
        out = struct.pack(">i", 1) + struct.pack(">i", 7) + enc_string("127.0.0.1") + struct.pack(">i", self.port) + enc_string(None) + struct.pack(">i", 7)
        out += struct.pack(">i", len(topics))
        for t in topics:
            parts = sorted(p for (tt, p) in self.logs if tt == t)
            out += struct.pack(">h", 0 if parts else 3) + enc_string(t) + struct.pack(">b", 0) + struct.pack(">i", len(parts))
            for p in parts:
                out += struct.pack(">hii", 0, p, 7) + struct.pack(">ii", 1, 7) + struct.pack(">ii", 1, 7)
        return out

    def _list_offsets(self, r):
        r.i32()
        out = b""
        nt = r.i32()
        out += struct.pack(">i", nt)
        for _ in range(nt):
            t = r.string(); npart = r.i32()
            out += enc_string(t) + struct.pack(">i", npart)
            for _ in range(npart):
                p, ts = r.i32(), r.i64()
                known = (t, p) in self.logs
                off = 0 if ts == -2 else self.next.get((t, p), 0)
                out += struct.pack(">ihqq", p, 0 if known else 3, -1, off)
        return out

    def _fetch(self, r):
        r.i32(); max_wait = r.i32(); r.i32(); r.i32(); r.i8()
        reqs = []
        for _ in range(r.i32()):
            t = r.string()
            for _ in range(r.i32()):
                p, off, pmax = r.i32(), r.i64(), r.i32()
                reqs.append((t, p, off, pmax))
        deadline = time.time() + max_wait / 1e3
        while True:
            with self.lock:
                have = any(self.next.get((t, p), 0) > off for t, p, off, _ in reqs)
            if have or time.time() >= deadline or self.fail_next_fetch_with:
                break
            time.sleep(0.005)
        out = struct.pack(">i", 0) + struct.pack(">i", len(reqs))
        for t, p, off, pmax in reqs:
            err, self.fail_next_fetch_with = self.fail_next_fetch_with, 0
            with self.lock:
                hw = self.next.get((t, p), 0)
                if (t, p) not in self.logs:
                    err = err or 3
                elif off > hw:
                    err = err or 1
                data = b""
                if not err:
                    for base, n, raw in self.logs[(t, p)]:
                        if base + n > off and len(data) + len(raw) <= max(pmax, len(raw)):
                            data += raw
                    if self.truncate_last and len(data) > 40:
                        data = data + self.logs[(t, p)][-1][2][:23]          # a partial batch trails the complete ones
            out += enc_string(t) + struct.pack(">i", 1) + struct.pack(">ihqq", p, err, hw, hw) + struct.pack(">i", -1) + struct.pack(">i", len(data)) + data
        return out
