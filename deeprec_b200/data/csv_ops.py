"""CSV / string helper ops (kernels/trans_csv_ali_ops.cc, string_split_and_pad_ali_op.cc, sparse_valid_cutoff_op.cc in the reference)."""
from __future__ import annotations

import zlib
from typing import List, Sequence

import torch

from ..ops.embedding_ops import SparseIds


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


def trans_csv_id2sparse(records: Sequence[str], max_id: int = -1, id_as_value: bool = True, field_delim: str = ",") -> SparseIds:
    """``TransCsvID2Sparse``: each record is a delimiter-separated list of ids -> SparseIds (ids > max_id dropped when max_id >= 0)."""
    vals, rows = [], []
    for r, rec in enumerate(records):
        for tok in rec.split(field_delim):
            tok = tok.strip()
            if not tok:
                continue
            v = int(tok)
            if max_id >= 0 and v > max_id:
                continue
            vals.append(v); rows.append(r)
    return SparseIds(torch.tensor(vals, dtype=torch.int64), torch.tensor(rows, dtype=torch.int64), len(records))


def trans_csv_kv2dense(records: Sequence[str], max_id: int, field_delim: str = ",", kv_delim: str = ":") -> torch.Tensor:
    """``TransCsvKV2Dense``: "k:v,k:v" records -> dense [N, max_id + 1]."""
    out = torch.zeros(len(records), max_id + 1)
    for r, rec in enumerate(records):
        for tok in rec.split(field_delim):
            if kv_delim in tok:
                k, v = tok.split(kv_delim, 1)
                k = int(k)
                if 0 <= k <= max_id:
                    out[r, k] = float(v)
    return out


def sparse_valid_cutoff(sp: SparseIds, cutoff_length: int, side: str = "right") -> SparseIds:
    """``SparseValidCutoff``: keep at most ``cutoff_length`` entries per row (from the left, or the right-most ones)."""
    counts = torch.bincount(sp.row_ids, minlength=sp.batch_size)
    starts = torch.cumsum(counts, 0) - counts
    idx_in_row = torch.arange(sp.values.numel()) - starts[sp.row_ids]
    keep = idx_in_row < cutoff_length if side == "left" else idx_in_row >= (counts[sp.row_ids] - cutoff_length)
    return SparseIds(sp.values[keep], sp.row_ids[keep], sp.batch_size, sp.weights[keep] if sp.weights is not None else None)
