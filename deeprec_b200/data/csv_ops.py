"""CSV / string helper ops of the input pipeline -- ``TransCsvID2Sparse / ID2Dense / KV2Sparse / KV2Dense / ToDense``, ``StringSplitAndPad``,
``SparseValidCutoff`` (kernels/trans_csv_ali_ops.cc, string_split_and_pad_ali_op.cc, sparse_valid_cutoff_op.cc in the reference).

The decoders are native (``csrc/host/csv_ops.cc``): the records of a batch are packed into one byte buffer and parsed in two parallel passes
(count, fill) on the host runtime's OpenMP pool.  Semantics follow the reference: the second dimension of a result is ``max_id`` (ids / keys live in
``[0, max_id)``; ``max_id=None`` detects it from the data, the reference's ``ID_AUTO_DETECT_TAG``), empty tokens are skipped, malformed numbers raise."""
from __future__ import annotations

import ctypes as C
import zlib
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _native
from ..ops.embedding_ops import SparseIds

Records = Union[Sequence[str], Sequence[bytes]]


def _lib():
    lib = _native.host()
    if not getattr(lib, "_csv_bound", False):
        P, i64, ch = C.c_void_p, C.c_int64, C.c_char
        lib.dr_csv_count.argtypes, lib.dr_csv_count.restype = [P, P, i64, ch, P], i64
        lib.dr_csv_ids.argtypes, lib.dr_csv_ids.restype = [P, P, i64, ch, i64, P, P, P], i64
        lib.dr_csv_kvs.argtypes, lib.dr_csv_kvs.restype = [P, P, i64, ch, ch, i64, P, P, P, P], i64
        lib.dr_csv_to_dense.argtypes, lib.dr_csv_to_dense.restype = [P, P, i64, ch, i64, P], i64
        lib.dr_csv_split_pad_ids.argtypes, lib.dr_csv_split_pad_ids.restype = [P, P, i64, ch, i64, i64, P], i64
        lib._csv_bound = True
    return lib


def _pack(records: Records) -> Tuple[np.ndarray, np.ndarray]:
    bs = [r if isinstance(r, bytes) else str(r).encode() for r in records]
    offs = np.zeros(len(bs) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in bs], out=offs[1:])
    buf = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8)
    return buf, offs


def _delim(d: str) -> bytes:
    if len(d) != 1 or d in "0123456789+-.eE ":
        raise ValueError(f"unsupported delimiter {d!r}")
    return d.encode()


def _counts(buf, offs, n, delim) -> Tuple[np.ndarray, int]:
    counts = np.zeros(n, dtype=np.int64)
    total = _lib().dr_csv_count(buf.ctypes.data, offs.ctypes.data, n, delim, counts.ctypes.data)
    return counts, int(total)


def string_to_hash_id(s: str) -> int:
    """Stable 63-bit id of a categorical string (the reference hashes strings to int64 keys for the EV)."""
    b = s.encode()
    return ((zlib.crc32(b) << 31) ^ zlib.adler32(b)) & 0x7FFFFFFFFFFFFFFF


def string_split_and_pad(strings: Sequence[str], max_length: int, delimiter: str = ",", default_value: str = "") -> List[List[str]]:
    """``StringSplitAndPad``: split every string, truncate / pad to ``max_length`` tokens."""
    out = []
    for s in strings:
        toks = s.split(delimiter) if s else []
        toks = toks[:max_length] + [default_value] * max(0, max_length - len(toks))
        out.append(toks)
    return out


def string_split_and_pad_ids(records: Records, max_length: int, delimiter: str = ",", pad_value: int = -1) -> torch.Tensor:
    """The same on id strings, natively: ``[N, max_length]`` int64, short records padded with ``pad_value`` (the ``[B, L]`` behaviour-history form)."""
    buf, offs = _pack(records)
    n = len(offs) - 1
    out = np.empty((n, int(max_length)), dtype=np.int64)
    bad = _lib().dr_csv_split_pad_ids(buf.ctypes.data, offs.ctypes.data, n, _delim(delimiter), int(max_length), int(pad_value), out.ctypes.data)
    if bad:
        raise ValueError(f"string_split_and_pad_ids: {bad} malformed tokens")
    return torch.from_numpy(out)


def _ids(records: Records, max_id: Optional[int], field_delim: str):
    buf, offs = _pack(records)
    n = len(offs) - 1
    d = _delim(field_delim)
    counts, total = _counts(buf, offs, n, d)
    start = np.zeros(n, dtype=np.int64); np.cumsum(counts[:-1], out=start[1:]) if n > 1 else None
    rows, ids = np.empty(total, dtype=np.int64), np.empty(total, dtype=np.int64)
    bad = _lib().dr_csv_ids(buf.ctypes.data, offs.ctypes.data, n, d, -1 if max_id is None else int(max_id), start.ctypes.data, rows.ctypes.data, ids.ctypes.data)
    if bad:
        raise ValueError(f"trans_csv: {bad} malformed id tokens")
    keep = ids >= 0
    if max_id is not None and not keep.all():
        raise ValueError(f"trans_csv: ids outside [0, {max_id})")
    return torch.from_numpy(rows), torch.from_numpy(ids), n


def trans_csv_id2sparse(records: Records, max_id: Optional[int] = None, id_as_value: bool = True, default_value: float = 1.0,
                        field_delim: str = ",") -> SparseIds:
    """``TransCsvID2Sparse``: ``["2,10", "7", "0,8"]`` -> entries (row, id); the entry's value is the id itself (``id_as_value``) or ``default_value``
    (carried as the SparseIds weight).  ``dense_shape = [N, max_id]`` (``.dense_shape``)."""
    rows, ids, n = _ids(records, max_id, field_delim)
    sp = SparseIds(ids, rows, n, None if id_as_value else torch.full((ids.numel(),), float(default_value)))
    sp.dense_shape = (n, int(max_id) if max_id is not None else (int(ids.max()) + 1 if ids.numel() else 0))
    return sp


def trans_csv_id2dense(records: Records, max_id: Optional[int] = None, id_as_value: bool = False, default_value: float = 1.0, field_delim: str = ",",
                       dtype: torch.dtype = torch.int64) -> torch.Tensor:
    """``TransCsvID2Dense``: multi-hot rows -- ``out[r, id] = default_value`` (or the id itself), 0 elsewhere; ``[N, max_id]``."""
    rows, ids, n = _ids(records, max_id, field_delim)
    width = int(max_id) if max_id is not None else (int(ids.max()) + 1 if ids.numel() else 0)
    out = torch.zeros(n, width, dtype=dtype)
    out[rows, ids] = ids.to(dtype) if id_as_value else torch.full((ids.numel(),), default_value, dtype=dtype)
    return out


def _kvs(records: Records, max_id: Optional[int], field_delim: str, kv_delim: str):
    if field_delim == kv_delim:
        raise ValueError("field and key-value delimiters must differ")
    buf, offs = _pack(records)
    n = len(offs) - 1
    d = _delim(field_delim)
    counts, total = _counts(buf, offs, n, d)
    start = np.zeros(n, dtype=np.int64); np.cumsum(counts[:-1], out=start[1:]) if n > 1 else None
    rows, keys, vals = np.empty(total, dtype=np.int64), np.empty(total, dtype=np.int64), np.empty(total, dtype=np.float32)
    bad = _lib().dr_csv_kvs(buf.ctypes.data, offs.ctypes.data, n, d, _delim(kv_delim), -1 if max_id is None else int(max_id), start.ctypes.data, rows.ctypes.data,
                            keys.ctypes.data, vals.ctypes.data)
    if bad:
        raise ValueError(f"trans_csv: {bad} malformed key:value tokens")
    if max_id is not None and (keys < 0).any():
        raise ValueError(f"trans_csv: keys outside [0, {max_id})")
    return torch.from_numpy(rows), torch.from_numpy(keys), torch.from_numpy(vals), n


def trans_csv_kv2sparse(records: Records, max_id: Optional[int] = None, field_delim: str = ",", kv_delim: str = ":") -> SparseIds:
    """``TransCsvKV2Sparse``: ``["2:2.0,10:0.1", "7:-0.7"]`` -> entries (row, key) weighted by the value."""
    rows, keys, vals, n = _kvs(records, max_id, field_delim, kv_delim)
    sp = SparseIds(keys, rows, n, vals)
    sp.dense_shape = (n, int(max_id) if max_id is not None else (int(keys.max()) + 1 if keys.numel() else 0))
    return sp


def trans_csv_kv2dense(records: Records, max_id: Optional[int] = None, field_delim: str = ",", kv_delim: str = ":") -> torch.Tensor:
    """``TransCsvKV2Dense``: ``["2:0.2,1:0.1", "3:-0.3"]`` with ``max_id = 4`` -> ``[[0, .1, .2, 0], [0, 0, 0, -.3]]``."""
    rows, keys, vals, n = _kvs(records, max_id, field_delim, kv_delim)
    width = int(max_id) if max_id is not None else (int(keys.max()) + 1 if keys.numel() else 0)
    out = torch.zeros(n, width)
    out[rows, keys] = vals
    return out


def trans_csv_to_dense(records: Records, max_id: Optional[int] = None, field_delim: str = ",") -> torch.Tensor:
    """``TransCsvToDense``: plain number rows, left-aligned and zero-padded to ``max_id`` columns (``None``: the longest record)."""
    buf, offs = _pack(records)
    n = len(offs) - 1
    d = _delim(field_delim)
    if max_id is None:
        counts, _ = _counts(buf, offs, n, d)
        max_id = int(counts.max()) if n else 0
    out = np.empty((n, int(max_id)), dtype=np.float32)
    bad = _lib().dr_csv_to_dense(buf.ctypes.data, offs.ctypes.data, n, d, int(max_id), out.ctypes.data)
    if bad:
        raise ValueError(f"trans_csv_to_dense: {bad} malformed numbers")
    return torch.from_numpy(out)


def sparse_valid_cutoff(sp: SparseIds, cutoff_length: int, side: str = "right") -> SparseIds:
    """``SparseValidCutoff``: keep at most ``cutoff_length`` entries per row (from the left, or the right-most ones)."""
    counts = torch.bincount(sp.row_ids, minlength=sp.batch_size)
    starts = torch.cumsum(counts, 0) - counts
    idx_in_row = torch.arange(sp.values.numel()) - starts[sp.row_ids]
    keep = idx_in_row < cutoff_length if side == "left" else idx_in_row >= (counts[sp.row_ids] - cutoff_length)
    return SparseIds(sp.values[keep], sp.row_ids[keep], sp.batch_size, sp.weights[keep] if sp.weights is not None else None)
