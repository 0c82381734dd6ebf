"""ParquetDataset / DataFrame: columnar batch reader -> dict of dense tensors / ragged values / SparseIds.

Parity: core/kernels/data/{parquet_dataset_ops,parquet_batch_reader,arrow_util}.cc + python/data/experimental/ops/
{parquet_dataset_ops,dataframe,parquet_pybind}.py (docs/docs_en/Parquet-Dataset.md):

* ``ParquetDataset(filenames, batch_size, fields=None, partition_count=1, partition_index=0, drop_remainder=False, num_parallel_reads=None,
  num_sequential_reads=1)`` -- ``filenames`` a path, a list, or ANY iterable of paths consumed lazily (``WorkQueue.input_dataset()``: workers pull
  files from the shared queue while they read);
* ``fields``: names or :class:`DataFrameField` (name, dtype, ragged_rank, shape) -- column projection happens in the Parquet reader, a declared
  dtype is cast on the way out, a declared fixed ``shape`` turns a list column into a dense ``[B, *shape]`` tensor;
* list columns of ANY nesting depth come back as :class:`DataFrameValue` (values + one row-splits vector per ragged level, outermost first),
  ``to_sparse()`` gives the ``SparseIds`` the embedding lookups take;
* every batch has exactly ``batch_size`` rows -- row groups and file boundaries are re-batched (the fused engines replay a CUDA graph with static
  shapes); the tail is dropped with ``drop_remainder`` or yielded short;
* ``partition_count / partition_index`` split the ROW GROUPS of every file among the data-parallel workers (no file-count constraint, no overlap);
* ``num_parallel_reads`` files are decoded concurrently by a thread pool (Arrow releases the GIL) and their batches interleaved ``num_sequential_reads``
  at a time, deterministically.

Arrow (pyarrow) does the decoding, as Arrow C++ does in the reference; ``parquet_fields`` reads the schema like ``parquet_pybind``."""
from __future__ import annotations

import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from ..ops.embedding_ops import SparseIds


@dataclass
class DataFrameValue:
    """``DataFrame.Value``: flat values + nested row splits of a ragged column (outermost level first, as ``tf.RaggedTensor``)."""
    values: torch.Tensor
    nested_row_splits: List[torch.Tensor]

    @property
    def ragged_rank(self) -> int:
        return len(self.nested_row_splits)

    def to_sparse(self) -> SparseIds:
        """Rows = the OUTERMOST dimension (one row per sample); deeper levels are flattened into the row (ids of all inner lists)."""
        splits = self.nested_row_splits[0]
        for inner in self.nested_row_splits[1:]:
            splits = inner[splits]
        return SparseIds.from_offsets(self.values, splits)

    def to_list(self) -> list:
        def build(level: int, lo: int, hi: int):
            if level == len(self.nested_row_splits):
                return self.values[lo:hi].tolist()
            sp = self.nested_row_splits[level]
            return [build(level + 1, int(sp[i]), int(sp[i + 1])) for i in range(lo, hi)]
        return build(0, 0, int(self.nested_row_splits[0].numel()) - 1)


@dataclass
class DataFrameField:
    """``DataFrame.Field``: one selected column -- its name, the dtype it is delivered in, how many ragged levels it has, and (for list columns whose
    lists all have the same length) a fixed inner ``shape`` that makes it a dense tensor."""
    name: str
    dtype: Optional[torch.dtype] = None
    ragged_rank: Optional[int] = None
    shape: Optional[Sequence[int]] = None
    incomplete: bool = field(default=False, repr=False)


_ARROW_TO_TORCH = {"int8": torch.int8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64, "uint8": torch.uint8, "float": torch.float32,
                   "halffloat": torch.float16, "double": torch.float64, "bool": torch.bool}


def _leaf_and_rank(t) -> Tuple[object, int]:
    import pyarrow as pa
    rank = 0
    while pa.types.is_list(t) or pa.types.is_large_list(t) or pa.types.is_fixed_size_list(t):
        t, rank = t.value_type, rank + 1
    return t, rank


def parquet_fields(filename: str, fields: Optional[Sequence[Union[str, DataFrameField]]] = None) -> List[DataFrameField]:
    """Schema of a file as ``DataFrameField``s (``parquet_pybind.parquet_fields``): dtype and ragged rank of every (selected) column; strings are
    delivered as int64 hashes."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    schema = pq.read_schema(filename)
    wanted = None if fields is None else {(f.name if isinstance(f, DataFrameField) else f): f for f in fields}
    if wanted is not None:
        missing = [n for n in wanted if n not in schema.names]
        if missing:
            raise ValueError(f"{filename}: no such columns {missing}; the file has {schema.names}")
    out = []
    for name in (schema.names if wanted is None else list(wanted)):
        leaf, rank = _leaf_and_rank(schema.field(name).type)
        dt = torch.int64 if (pa.types.is_string(leaf) or pa.types.is_large_string(leaf) or pa.types.is_binary(leaf)) else _ARROW_TO_TORCH.get(str(leaf))
        if dt is None:
            raise TypeError(f"{filename}: column {name} has unsupported type {leaf}")
        decl = wanted[name] if wanted is not None and isinstance(wanted[name], DataFrameField) else None
        if decl is not None and decl.ragged_rank is not None and decl.ragged_rank != rank and decl.shape is None:
            raise ValueError(f"{filename}: column {name} has ragged rank {rank}, declared {decl.ragged_rank}")
        out.append(DataFrameField(name, (decl.dtype if decl is not None and decl.dtype is not None else dt), rank, decl.shape if decl is not None else None))
    return out


def _hash_strings(a: np.ndarray) -> np.ndarray:
    """strings -> stable 63-bit ids (categorical features arrive hashed, as after ``tf.strings.to_hash_bucket_fast`` without the bucket)."""
    return np.fromiter((((zlib.crc32(s) << 31) ^ zlib.adler32(s)) & 0x7FFFFFFFFFFFFFFF for s in (x if isinstance(x, bytes) else str(x).encode() for x in a)),
                       dtype=np.int64, count=len(a))


def _leaf_tensor(arr, dtype: Optional[torch.dtype]) -> torch.Tensor:
    import pyarrow as pa
    if pa.types.is_string(arr.type) or pa.types.is_large_string(arr.type) or pa.types.is_binary(arr.type):
        a = _hash_strings(arr.to_numpy(zero_copy_only=False))
    else:
        if arr.null_count:
            arr = arr.fill_null(0)
        a = arr.to_numpy(zero_copy_only=False)
    a = np.ascontiguousarray(a)
    t = torch.from_numpy(a if a.flags.writeable else a.copy())
    return t.to(dtype) if dtype is not None and t.dtype != dtype else t


def _column(col, fld: Optional[DataFrameField]) -> Union[torch.Tensor, DataFrameValue]:
    import pyarrow as pa
    arr = col.combine_chunks() if isinstance(col, pa.ChunkedArray) else col
    splits: List[torch.Tensor] = []
    n_rows = len(arr)
    while pa.types.is_list(arr.type) or pa.types.is_large_list(arr.type) or pa.types.is_fixed_size_list(arr.type):
        if pa.types.is_fixed_size_list(arr.type):
            k = arr.type.list_size
            offs = np.arange(len(arr) + 1, dtype=np.int64) * k
            arr = arr.flatten()
        else:
            offs = np.asarray(arr.offsets, dtype=np.int64)
            lo, hi = int(offs[0]), int(offs[-1])
            arr = arr.values.slice(lo, hi - lo)                    # a sliced list array keeps its parent's value buffer
            offs = offs - lo
        splits.append(torch.from_numpy(np.ascontiguousarray(offs)))
    vals = _leaf_tensor(arr, fld.dtype if fld is not None else None)
    if not splits:
        return vals
    if fld is not None and fld.shape is not None:                  # declared fixed inner shape: a dense [B, *shape] tensor
        want = 1
        for d in fld.shape:
            want *= int(d)
        if vals.numel() != n_rows * want:
            raise ValueError(f"column {fld.name}: declared shape {list(fld.shape)} but {vals.numel()} values for {n_rows} rows")
        return vals.view(n_rows, *[int(d) for d in fld.shape])
    return DataFrameValue(vals, splits)


class ParquetDataset:
    def __init__(self, filenames: Union[str, Sequence[str], Iterable[str]], batch_size: int = 1024,
                 fields: Optional[Sequence[Union[str, DataFrameField]]] = None, partition_count: int = 1, partition_index: int = 0,
                 drop_remainder: bool = False, num_epochs: int = 1, num_parallel_reads: Optional[int] = None, num_sequential_reads: int = 1,
                 rebatch_across_files: bool = True):
        import pyarrow.parquet as pq  # noqa: F401  (fail at construction, not at the first batch)
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        if not (0 <= partition_index < max(1, partition_count)):
            raise ValueError(f"partition_index {partition_index} outside [0, {partition_count})")
        self._lazy = not isinstance(filenames, (str, list, tuple))
        self.files = [filenames] if isinstance(filenames, str) else (filenames if self._lazy else list(filenames))
        self.batch_size, self.drop_remainder, self.num_epochs = int(batch_size), bool(drop_remainder), int(num_epochs)
        self.partition_count, self.partition_index = int(partition_count), int(partition_index)
        self.num_parallel_reads = int(num_parallel_reads) if num_parallel_reads else 1
        self.num_sequential_reads = max(1, int(num_sequential_reads))
        self.rebatch_across_files = bool(rebatch_across_files)
        self._decl = None if fields is None else [f if isinstance(f, DataFrameField) else DataFrameField(str(f)) for f in fields]
        self.fields = None if self._decl is None else [f.name for f in self._decl]

    # ---- one file -> stream of Arrow tables (its row groups of this partition) ----------------------------------------------------------------
    def _tables_of(self, path: str):
        import pyarrow.parquet as pq
        pf = pq.ParquetFile(path)
        if self._decl is not None:
            parquet_fields(path, self._decl)                         # validates names / declared ragged ranks against the schema
        groups = [g for g in range(pf.num_row_groups) if g % self.partition_count == self.partition_index]
        for g in groups:
            yield pf.read_row_group(g, columns=self.fields)

    def _file_stream(self, files: Iterable[str]):
        """Arrow tables in file order; with num_parallel_reads > 1 a window of files is decoded concurrently and interleaved
        ``num_sequential_reads`` row groups at a time (deterministic for a given file order)."""
        if self.num_parallel_reads <= 1:
            for f in files:
                yield from self._tables_of(f)
            return
        it = iter(files)
        with ThreadPoolExecutor(max_workers=self.num_parallel_reads) as ex:
            while True:
                window = []
                for _ in range(self.num_parallel_reads):
                    f = next(it, None)
                    if f is None:
                        break
                    window.append(f)
                if not window:
                    return
                decoded = list(ex.map(lambda p: list(self._tables_of(p)), window))
                pos = [0] * len(decoded)
                while any(pos[i] < len(decoded[i]) for i in range(len(decoded))):
                    for i in range(len(decoded)):
                        for _ in range(self.num_sequential_reads):
                            if pos[i] < len(decoded[i]):
                                yield decoded[i][pos[i]]; pos[i] += 1

    def _emit(self, table) -> Dict[str, Union[torch.Tensor, DataFrameValue]]:
        decl = {f.name: f for f in self._decl} if self._decl is not None else {}
        return {name: _column(table.column(i), decl.get(name)) for i, name in enumerate(table.schema.names)}

    def __iter__(self) -> Iterator[Dict[str, Union[torch.Tensor, DataFrameValue]]]:
        import pyarrow as pa
        for _ in range(self.num_epochs):
            pending, rows = [], 0
            for tbl in self._file_stream(self.files):
                pending.append(tbl); rows += tbl.num_rows
                if rows < self.batch_size:
                    continue
                big = pa.concat_tables(pending) if len(pending) > 1 else pending[0]
                off = 0
                while rows - off >= self.batch_size:
                    yield self._emit(big.slice(off, self.batch_size)); off += self.batch_size
                pending, rows = ([big.slice(off)] if rows - off else []), rows - off
            if rows and not self.drop_remainder:
                yield self._emit(pa.concat_tables(pending) if len(pending) > 1 else pending[0])
            if self._lazy:                                             # a lazily consumed source (a work queue) is one pass by definition
                return


def read_parquet(filenames, batch_size=1024, fields=None, **kw) -> ParquetDataset:
    return ParquetDataset(filenames, batch_size, fields, **kw)
