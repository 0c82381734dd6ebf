"""Feature stores for serving (serving/processor/storage/{feature_store.h, feature_store_mgr.h, redis_feature_store.*} in the
reference; ``feature_store_type: local | redis`` in the model config).

* ``LocalFeatureStore``   -- rows live in this process (in-process EmbeddingVariables; the default).
* ``RedisFeatureStore``   -- rows live in a Redis instance shared by every serving replica.  The client is native
  (``csrc/host/redis_store.cc``: RESP2 over TCP, pipelined MGET / MSET) and a small connection pool gives each serving session its
  own socket.  ``MiniRedisServer`` is an in-process RESP server (tests, single-box deployments without a Redis binary).
* ``export_to_feature_store`` = the reference's full-model import (KvImport): snapshot every EmbeddingVariable into the store under
  ``<model>/<version>/<table>`` and publish the version key last, so readers never see a half-written version.
* ``attach_feature_store`` swaps a model's EmbeddingVariable lookups for store lookups (the SavedModelOptimizer's
  EV-op -> KvLookup rewrite); rows the store does not have fall back to the variable's default row, exactly like a local lookup.
"""
from __future__ import annotations

import ctypes as C
import queue
import socketserver
import threading
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _native

_BOUND = False


def _lib():
    global _BOUND
    L = _native.host()
    if not _BOUND:
        vp, i64, cp = C.c_void_p, C.c_int64, C.c_char_p
        L.dr_redis_connect.restype, L.dr_redis_connect.argtypes = vp, [cp, C.c_int, C.c_int, cp, C.c_int]
        L.dr_redis_ok.restype, L.dr_redis_ok.argtypes = C.c_int, [vp]
        L.dr_redis_last_error.restype, L.dr_redis_last_error.argtypes = cp, [vp]
        L.dr_redis_close.argtypes = [vp]
        L.dr_redis_ping.restype, L.dr_redis_ping.argtypes = C.c_int, [vp]
        L.dr_redis_dbsize.restype, L.dr_redis_dbsize.argtypes = i64, [vp]
        L.dr_redis_flushdb.restype, L.dr_redis_flushdb.argtypes = C.c_int, [vp]
        L.dr_redis_set.restype, L.dr_redis_set.argtypes = C.c_int, [vp, cp, cp, i64]
        L.dr_redis_get.restype, L.dr_redis_get.argtypes = i64, [vp, cp, vp, i64]
        L.dr_redis_mset_rows.restype, L.dr_redis_mset_rows.argtypes = i64, [vp, cp, vp, i64, vp, C.c_int]
        L.dr_redis_mget_rows.restype, L.dr_redis_mget_rows.argtypes = i64, [vp, cp, vp, i64, vp, C.c_int, vp]
        L.dr_redis_del_rows.restype, L.dr_redis_del_rows.argtypes = i64, [vp, cp, vp, i64]
        _BOUND = True
    return L


class FeatureStore:
    """``IFeatureStoreMgr``: batched row lookup / insert per table + model-version bookkeeping."""

    def lookup(self, table: str, keys: np.ndarray, dim: int) -> Tuple[np.ndarray, np.ndarray]:
        """-> (rows [n, dim] float32 -- undefined where ``found`` is False, found [n] bool)"""
        raise NotImplementedError

    def insert(self, table: str, keys: np.ndarray, rows: np.ndarray) -> None:
        raise NotImplementedError

    def remove(self, table: str, keys: np.ndarray) -> int:
        raise NotImplementedError

    def set_meta(self, key: str, value: str) -> None:
        raise NotImplementedError

    def get_meta(self, key: str) -> Optional[str]:
        raise NotImplementedError

    def close(self) -> None:
        pass


class LocalFeatureStore(FeatureStore):
    """In-process store (numpy rows behind a dict index)."""

    def __init__(self):
        self._tables: Dict[str, Dict[int, np.ndarray]] = {}
        self._meta: Dict[str, str] = {}
        self._mu = threading.Lock()

    def lookup(self, table, keys, dim):
        keys = np.asarray(keys, dtype=np.int64).reshape(-1)
        rows, found = np.zeros((keys.size, dim), np.float32), np.zeros(keys.size, bool)
        t = self._tables.get(table, {})
        for i, k in enumerate(keys.tolist()):
            r = t.get(k)
            if r is not None:
                rows[i], found[i] = r, True
        return rows, found

    def insert(self, table, keys, rows):
        rows = np.asarray(rows, dtype=np.float32)
        with self._mu:
            t = self._tables.setdefault(table, {})
            for k, r in zip(np.asarray(keys, dtype=np.int64).reshape(-1).tolist(), rows):
                t[k] = r.copy()

    def remove(self, table, keys):
        with self._mu:
            t = self._tables.get(table, {})
            return sum(t.pop(k, None) is not None for k in np.asarray(keys, dtype=np.int64).reshape(-1).tolist())

    def set_meta(self, key, value):
        self._meta[key] = value

    def get_meta(self, key):
        return self._meta.get(key)


class RedisFeatureStore(FeatureStore):
    """Rows in Redis: key ``<table>:<id>``, value = dim x fp32.  ``pool_size`` native connections (one per concurrent caller)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 6379, password: str = "", db: int = 0, timeout_ms: int = 5000, pool_size: int = 4):
        self.L = _lib()
        self._pool: "queue.Queue[int]" = queue.Queue()
        self._all: List[int] = []
        for _ in range(max(1, pool_size)):
            h = self.L.dr_redis_connect(host.encode(), port, timeout_ms, password.encode(), db)
            if not self.L.dr_redis_ok(h):
                err = self.L.dr_redis_last_error(h).decode()
                self.L.dr_redis_close(h)
                self.close()
                raise ConnectionError(f"redis feature store {host}:{port}: {err}")
            self._all.append(h)
            self._pool.put(h)

    class _Lease:
        def __init__(self, store):
            self.s = store

        def __enter__(self):
            self.h = self.s._pool.get()
            return self.h

        def __exit__(self, *exc):
            self.s._pool.put(self.h)

    def _err(self, h) -> str:
        return self.L.dr_redis_last_error(h).decode()

    def ping(self) -> bool:
        with self._Lease(self) as h:
            return self.L.dr_redis_ping(h) == 0

    def dbsize(self) -> int:
        with self._Lease(self) as h:
            return int(self.L.dr_redis_dbsize(h))

    def flush(self) -> None:
        with self._Lease(self) as h:
            self.L.dr_redis_flushdb(h)

    def lookup(self, table, keys, dim):
        keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        rows, found = np.zeros((keys.size, dim), np.float32), np.zeros(keys.size, np.uint8)
        if keys.size:
            with self._Lease(self) as h:
                if self.L.dr_redis_mget_rows(h, table.encode(), keys.ctypes.data, keys.size, rows.ctypes.data, dim, found.ctypes.data) < 0:
                    raise IOError(f"redis MGET failed: {self._err(h)}")
        return rows, found.astype(bool)

    def insert(self, table, keys, rows):
        keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(keys.size, -1)
        if keys.size:
            with self._Lease(self) as h:
                if self.L.dr_redis_mset_rows(h, table.encode(), keys.ctypes.data, keys.size, rows.ctypes.data, rows.shape[1]) < 0:
                    raise IOError(f"redis MSET failed: {self._err(h)}")

    def remove(self, table, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
        with self._Lease(self) as h:
            return int(self.L.dr_redis_del_rows(h, table.encode(), keys.ctypes.data, keys.size))

    def set_meta(self, key, value):
        v = value.encode()
        with self._Lease(self) as h:
            if self.L.dr_redis_set(h, key.encode(), v, len(v)) != 0:
                raise IOError(f"redis SET failed: {self._err(h)}")

    def get_meta(self, key):
        buf = C.create_string_buffer(4096)
        with self._Lease(self) as h:
            n = self.L.dr_redis_get(h, key.encode(), buf, 4096)
        return None if n < 0 else buf.raw[:n].decode()

    def close(self):
        for h in self._all:
            self.L.dr_redis_close(h)
        self._all.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ in-process RESP server
class _RespHandler(socketserver.StreamRequestHandler):
    def _read_command(self) -> Optional[List[bytes]]:
        line = self.rfile.readline()
        if not line:
            return None
        if not line.startswith(b"*"):
            return line.split()                       # inline command
        args = []
        for _ in range(int(line[1:])):
            n = int(self.rfile.readline()[1:])
            args.append(self.rfile.read(n + 2)[:-2])
        return args

    def handle(self):
        srv: "MiniRedisServer" = self.server.owner       # type: ignore[attr-defined]
        with srv.mu:
            srv.conns.add(self.request)
        try:
            self._serve(srv)
        finally:
            with srv.mu:
                srv.conns.discard(self.request)

    def _serve(self, srv: "MiniRedisServer"):
        w = self.wfile
        authed = not srv.password
        while True:
            try:
                args = self._read_command()
            except (ConnectionError, ValueError, OSError):
                return
            if not args:
                return
            cmd = args[0].upper()
            if cmd == b"AUTH":
                authed = args[-1].decode() == srv.password
                w.write(b"+OK\r\n" if authed else b"-ERR invalid password\r\n")
            elif not authed:
                w.write(b"-NOAUTH Authentication required.\r\n")
            elif cmd == b"PING":
                w.write(b"+PONG\r\n")
            elif cmd == b"SELECT":
                w.write(b"+OK\r\n")
            elif cmd in (b"SET", b"MSET"):
                with srv.mu:
                    for i in range(1, len(args) - 1, 2):
                        srv.data[args[i]] = args[i + 1]
                w.write(b"+OK\r\n")
            elif cmd == b"GET":
                v = srv.data.get(args[1])
                w.write(b"$-1\r\n" if v is None else b"$%d\r\n%s\r\n" % (len(v), v))
            elif cmd == b"MGET":
                out = [b"*%d\r\n" % (len(args) - 1)]
                for k in args[1:]:
                    v = srv.data.get(k)
                    out.append(b"$-1\r\n" if v is None else b"$%d\r\n%s\r\n" % (len(v), v))
                w.write(b"".join(out))
            elif cmd == b"DEL":
                with srv.mu:
                    n = sum(srv.data.pop(k, None) is not None for k in args[1:])
                w.write(b":%d\r\n" % n)
            elif cmd == b"DBSIZE":
                w.write(b":%d\r\n" % len(srv.data))
            elif cmd == b"FLUSHDB":
                with srv.mu:
                    srv.data.clear()
                w.write(b"+OK\r\n")
            else:
                w.write(b"-ERR unknown command '%s'\r\n" % args[0])
            srv.commands += 1
            w.flush()


class MiniRedisServer:
    """Threaded RESP2 server with the command subset the feature store uses (PING AUTH SELECT GET SET MGET MSET DEL DBSIZE FLUSHDB)."""

    def __init__(self, host: str = "127.0.0.1", port: int = 0, password: str = ""):
        self.data: Dict[bytes, bytes] = {}
        self.mu = threading.Lock()
        self.password = password
        self.commands = 0
        self.conns = set()                            # established client sockets (closed by close(): a stopped server drops its clients)

        class _Srv(socketserver.ThreadingTCPServer):
            allow_reuse_address = True
            daemon_threads = True

        self._srv = _Srv((host, port), _RespHandler)
        self._srv.owner = self                       # type: ignore[attr-defined]
        self.host, self.port = self._srv.server_address[:2]
        self._thread = threading.Thread(target=self._srv.serve_forever, name="mini-redis", daemon=True)
        self._thread.start()

    def close(self):
        self._srv.shutdown()
        self._srv.server_close()
        import socket
        with self.mu:
            conns, self.conns = list(self.conns), set()
        for c in conns:
            try:
                c.shutdown(socket.SHUT_RDWR)
            except OSError:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


# ------------------------------------------------------------------------------------------------ model integration
def _table_prefix(model_name: str, version: int, table: str) -> str:
    return f"{model_name}/{version}/{table}"


def export_to_feature_store(embedding_variables: Iterable, store: FeatureStore, model_name: str, version: int, chunk: int = 65536) -> int:
    """Full import of every EmbeddingVariable's admitted rows (reference: KvImport on full model update).  The version key is
    written last.  Returns the number of rows written."""
    total = 0
    for ev in embedding_variables:
        keys, values, _, _ = ev.export()
        keys, values = keys.cpu().numpy(), values.cpu().numpy()
        pre = _table_prefix(model_name, version, ev.name)
        for o in range(0, keys.size, chunk):
            store.insert(pre, keys[o:o + chunk], values[o:o + chunk])
        total += int(keys.size)
    store.set_meta(f"{model_name}/latest_version", str(version))
    return total


def export_delta_to_feature_store(embedding_variables: Iterable, store: FeatureStore, model_name: str, version: int) -> int:
    """Delta update (KvInsert of the rows touched since the last save) into the CURRENT version's key space."""
    total = 0
    for ev in embedding_variables:
        s = ev.table.snapshot(dirty_only=True)
        keys = s["keys"].cpu().numpy()
        if keys.size:
            store.insert(_table_prefix(model_name, version, ev.name), keys, s["rows"][:, : ev.embedding_dim].contiguous().cpu().numpy())
            total += int(keys.size)
    return total


def export_processor_tables(embedding_variables: Iterable, store: FeatureStore, prefix: str, version: int, dirty_only: bool = False, chunk: int = 65536) -> int:
    """Rows for the NATIVE CPU Processor in ``feature_store_type: "redis"`` mode (csrc/host/cpu_serving.cc): table t of the exported model
    lives under ``<prefix>/<version>/table/<t>`` (``prefix`` = the processor's ``redis_prefix``).  Call it BEFORE publishing the
    saved-model version the rows belong to; ``dirty_only=True`` sends just the rows touched since the last export (delta update)."""
    total = 0
    for t, ev in enumerate(embedding_variables):
        s = ev.table.snapshot(dirty_only=dirty_only)
        keys, values = s["keys"].cpu().numpy(), s["rows"][:, : ev.embedding_dim].contiguous().cpu().numpy()
        for o in range(0, keys.size, chunk):
            store.insert(f"{prefix}/{int(version)}/table/{t}", keys[o:o + chunk], values[o:o + chunk])
        total += int(keys.size)
    return total


class FeatureStoreEmbedding(nn.Module):
    """Drop-in for an EmbeddingVariable at serving time: rows come from the store; unknown keys read the default row
    ``default_matrix[key % default_value_dim]`` (what a local inference-mode lookup returns)."""

    def __init__(self, store: FeatureStore, model_name: str, version: int, ev):
        super().__init__()
        self.store, self.name, self.embedding_dim = store, ev.name, ev.embedding_dim
        self.prefix = _table_prefix(model_name, version, ev.name)
        self.device = torch.device("cpu")
        self.register_buffer("default_matrix", ev.default_matrix.detach().clone().float().cpu())

    def set_version(self, model_name: str, version: int) -> None:
        self.prefix = _table_prefix(model_name, version, self.name)

    @torch.no_grad()
    def lookup(self, ids: torch.Tensor) -> torch.Tensor:
        flat = ids.reshape(-1).to(torch.int64).cpu()
        uniq, inv = torch.unique(flat, return_inverse=True)               # one store access per distinct key
        rows, found = self.store.lookup(self.prefix, uniq.numpy(), self.embedding_dim)
        rows_t, found_t = torch.from_numpy(rows), torch.from_numpy(found)
        if not bool(found_t.all()):
            dflt = self.default_matrix[torch.remainder(uniq, self.default_matrix.shape[0])]
            rows_t = torch.where(found_t[:, None], rows_t, dflt)
        return rows_t[inv].view(*ids.shape, self.embedding_dim)

    forward = lookup
    sparse_read = lookup


def attach_feature_store(model: nn.Module, store: FeatureStore, model_name: str, version: Optional[int] = None) -> List[FeatureStoreEmbedding]:
    """Replace every EmbeddingVariable submodule of ``model`` by a ``FeatureStoreEmbedding`` reading ``version`` (default: the
    store's published latest version)."""
    from ..embedding_variable import EmbeddingVariable
    if version is None:
        v = store.get_meta(f"{model_name}/latest_version")
        if v is None:
            raise LookupError(f"feature store has no published version of {model_name!r}")
        version = int(v)
    swapped: List[FeatureStoreEmbedding] = []

    def visit(mod: nn.Module):
        for name, child in list(mod.named_children()):
            if isinstance(child, EmbeddingVariable):
                fse = FeatureStoreEmbedding(store, model_name, version, child)
                if isinstance(mod, nn.ModuleList):
                    mod[int(name)] = fse
                else:
                    setattr(mod, name, fse)
                swapped.append(fse)
            else:
                visit(child)
    visit(model)
    return swapped
