"""ParquetDataset / DataFrame (data/parquet_dataset.py): exact re-batching across row groups and files, row-group partitioning among workers,
nested ragged columns, declared fields (dtype / fixed shape), schema introspection, parallel file reads, lazily consumed file sources (WorkQueue).

Reference: core/kernels/data/parquet_dataset_ops.cc, parquet_batch_reader.cc, python/data/experimental/ops/{parquet_dataset_ops,dataframe}.py,
docs/docs_en/Parquet-Dataset.md."""
import pytest
import torch

pa = pytest.importorskip("pyarrow")
import pyarrow.parquet as pq  # noqa: E402

from deeprec_b200.data import DataFrameField, ParquetDataset, WorkQueue, parquet_fields  # noqa: E402


def _write(path, lo, n, row_group_size):
    rows = list(range(lo, lo + n))
    tbl = pa.table({
        "id": pa.array(rows, pa.int64()),
        "f": pa.array([float(i) * 0.5 for i in rows], pa.float32()),
        "hist": pa.array([[i * 10 + j for j in range(i % 4)] for i in rows], pa.list_(pa.int64())),
        "nested": pa.array([[[i, i + 1][: (i + j) % 3] for j in range(i % 3)] for i in rows], pa.list_(pa.list_(pa.int32()))),
        "fixed": pa.array([[i, -i, 2 * i] for i in rows], pa.list_(pa.int64(), 3)),
        "tag": pa.array([f"t{i % 5}" for i in rows]),
    })
    pq.write_table(tbl, path, row_group_size=row_group_size)


def test_exact_rebatching_across_row_groups_and_files(tmp_path):
    files = []
    for k, n in enumerate((70, 45, 13)):                              # 128 rows in 3 files with row groups of 32 / 20 / 13 rows
        p = str(tmp_path / f"part{k}.parquet"); _write(p, 1000 * k, n, (32, 20, 13)[k]); files.append(p)
    ds = ParquetDataset(files, batch_size=25)
    batches = list(ds)
    assert [b["id"].numel() for b in batches] == [25, 25, 25, 25, 25, 3]        # every batch full except the tail
    ids = torch.cat([b["id"] for b in batches]).tolist()
    assert ids == list(range(0, 70)) + list(range(1000, 1045)) + list(range(2000, 2013))      # order preserved, nothing lost or repeated
    assert [b["id"].numel() for b in ParquetDataset(files, batch_size=25, drop_remainder=True)] == [25] * 5
    # ragged columns stay aligned with their rows through the re-batching
    for b in batches:
        h = b["hist"]
        assert h.ragged_rank == 1 and h.nested_row_splits[0].numel() == b["id"].numel() + 1
        assert h.to_list() == [[i * 10 + j for j in range(i % 4)] for i in b["id"].tolist()]
    assert list(ParquetDataset(files, batch_size=25, num_epochs=2)).__len__() == 12


def test_nested_ragged_declared_fields_and_schema(tmp_path):
    p = str(tmp_path / "d.parquet"); _write(p, 0, 40, 16)
    sch = {f.name: f for f in parquet_fields(p)}
    assert sch["hist"].ragged_rank == 1 and sch["nested"].ragged_rank == 2 and sch["id"].ragged_rank == 0 and sch["tag"].dtype == torch.int64
    assert sch["f"].dtype == torch.float32 and sch["nested"].dtype == torch.int32
    ds = ParquetDataset(p, batch_size=40, fields=[DataFrameField("id", torch.int32), DataFrameField("fixed", torch.int64, shape=[3]), "nested",
                                                   DataFrameField("f", torch.float64), "tag"])
    (b,) = list(ds)
    assert set(b) == {"id", "fixed", "nested", "f", "tag"} and b["id"].dtype == torch.int32 and b["f"].dtype == torch.float64
    assert b["fixed"].shape == (40, 3) and b["fixed"][7].tolist() == [7, -7, 14]                    # fixed-size lists -> a dense tensor
    nz = b["nested"]
    assert nz.ragged_rank == 2 and nz.to_list() == [[[i, i + 1][: (i + j) % 3] for j in range(i % 3)] for i in range(40)]
    sp = nz.to_sparse()                                                                          # inner lists flattened into the sample's row
    assert sp.batch_size == 40 and sp.values.numel() == nz.values.numel()
    assert torch.equal(torch.bincount(sp.row_ids, minlength=40), torch.tensor([sum(len([i, i + 1][: (i + j) % 3]) for j in range(i % 3)) for i in range(40)]))
    assert b["tag"].dtype == torch.int64 and b["tag"][0] == b["tag"][5] and b["tag"][0] != b["tag"][1]  # stable string hashes
    with pytest.raises(ValueError):
        list(ParquetDataset(p, batch_size=8, fields=["no_such_column"]))
    with pytest.raises(ValueError):
        list(ParquetDataset(p, batch_size=8, fields=[DataFrameField("hist", ragged_rank=2)]))


def test_row_group_partitioning_parallel_reads_and_work_queue_source(tmp_path):
    files = []
    for k in range(4):
        p = str(tmp_path / f"p{k}.parquet"); _write(p, 100 * k, 40, 10); files.append(p)       # 4 files x 4 row groups of 10 rows
    parts = [torch.cat([b["id"] for b in ParquetDataset(files, batch_size=16, partition_count=3, partition_index=r)]).tolist() for r in range(3)]
    allrows = sorted(x for part in parts for x in part)
    assert allrows == sorted(i for k in range(4) for i in range(100 * k, 100 * k + 40))             # disjoint, complete
    assert all(len(part) > 0 for part in parts)
    # parallel decoding: same multiset, deterministic order for a given file order
    a = torch.cat([b["id"] for b in ParquetDataset(files, batch_size=32, num_parallel_reads=3, num_sequential_reads=2)]).tolist()
    b2 = torch.cat([b["id"] for b in ParquetDataset(files, batch_size=32, num_parallel_reads=3, num_sequential_reads=2)]).tolist()
    assert a == b2 and sorted(a) == allrows
    # a lazily consumed source: workers take files from the shared queue while they read
    q = WorkQueue(files, num_epochs=1, shuffle=False)
    got = torch.cat([b["id"] for b in ParquetDataset(q.input_producer(), batch_size=64)]).tolist()
    assert sorted(got) == allrows
