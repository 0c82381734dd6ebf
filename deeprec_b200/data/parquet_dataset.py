"""ParquetDataset / DataFrame: columnar batch reader -> dict of dense tensors / SparseIds.

Parity: core/kernels/data/parquet_dataset_ops.cc + parquet_batch_reader.cc + python/data/experimental/ops/
{parquet_dataset_ops,dataframe}.py -- Arrow-based batches, field selection, ragged (list) columns returned as
values + row-splits (``DataFrame.Value``), partial last batch control.  Arrow (pyarrow) does the decoding, as Arrow C++ does
in the reference."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Union

import numpy as np
import torch

from ..ops.embedding_ops import SparseIds


@dataclass
class DataFrameValue:
    """``DataFrame.Value``: values + nested row splits of a ragged column."""
    values: torch.Tensor
    nested_row_splits: List[torch.Tensor]

    def to_sparse(self) -> SparseIds:
        return SparseIds.from_offsets(self.values, self.nested_row_splits[-1])


class ParquetDataset:
    def __init__(self, filenames: Union[str, Sequence[str]], batch_size: int = 1024, fields: Optional[Sequence[str]] = None,
                 partition_count: int = 1, partition_index: int = 0, drop_remainder: bool = False, num_epochs: int = 1):
        import pyarrow.parquet as pq  # noqa: F401
        self.files = [filenames] if isinstance(filenames, str) else list(filenames)
        self.batch_size, self.fields, self.drop_remainder, self.num_epochs = batch_size, list(fields) if fields else None, drop_remainder, num_epochs
        self.partition_count, self.partition_index = partition_count, partition_index

    @staticmethod
    def _column(col) -> Union[torch.Tensor, DataFrameValue]:
        import pyarrow as pa
        if pa.types.is_list(col.type) or pa.types.is_large_list(col.type):
            arr = col.combine_chunks() if hasattr(col, "combine_chunks") else col
            offs = torch.from_numpy(np.asarray(arr.offsets, dtype=np.int64).copy())
            vals = torch.from_numpy(np.asarray(arr.values).copy())
            return DataFrameValue(vals, [offs - offs[0]])
        a = col.to_numpy(zero_copy_only=False)
        if a.dtype == object:           # strings -> stable 63-bit hashes (categorical ids)
            import zlib
            a = np.fromiter(((zlib.crc32(str(x).encode()) << 31) ^ zlib.adler32(str(x).encode()) for x in a), dtype=np.int64, count=len(a))
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a if a.flags.writeable else a.copy())

    def __iter__(self) -> Iterator[Dict[str, Union[torch.Tensor, DataFrameValue]]]:
        import pyarrow.parquet as pq
        for _ in range(self.num_epochs):
            for f in self.files:
                pf = pq.ParquetFile(f)
                for gi, rb in enumerate(pf.iter_batches(batch_size=self.batch_size, columns=self.fields)):
                    if self.partition_count > 1 and gi % self.partition_count != self.partition_index:
                        continue
                    if self.drop_remainder and rb.num_rows < self.batch_size:
                        continue
                    yield {name: self._column(rb.column(i)) for i, name in enumerate(rb.schema.names)}


def read_parquet(filenames, batch_size=1024, fields=None, **kw) -> ParquetDataset:
    return ParquetDataset(filenames, batch_size, fields, **kw)
