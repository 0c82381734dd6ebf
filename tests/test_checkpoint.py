"""Full / incremental checkpoint round trips, N->M re-shard, filtered keys, eviction-at-save (incr_ckpt_test.py analogue)."""
import os

import pytest
import torch
from torch import nn

import deeprec_b200 as dr
from deeprec_b200.checkpoint import IncrementalSaver, Saver, latest_checkpoint
from deeprec_b200.optim import GlobalStep


class Tiny(nn.Module):
    def __init__(self, tag, **kw):
        super().__init__()
        self.ev = dr.get_embedding_variable(f"{tag}/emb", 8, ev_option=dr.EmbeddingVariableOption(**kw), seed=5)
        self.fc = nn.Linear(8, 1)

    def forward(self, ids):
        return self.fc(self.ev.lookup(ids)).squeeze(-1)


def _train(m, opt, steps, lo=0, hi=40):
    g = torch.Generator().manual_seed(steps + lo)
    for _ in range(steps):
        ids = torch.randint(lo, hi, (32,), generator=g)
        loss = (m(ids) - 1.0).pow(2).mean()
        opt.zero_grad(); loss.backward(); opt.step()


def test_full_roundtrip_with_slots_filter_and_optimizer_state(tmp_path):
    m = Tiny("ck1", filter_option=dr.CounterFilter(2))
    opt = dr.optim.AdamOptimizer(m, lr=0.01, global_step=GlobalStep())
    _train(m, opt, 6)
    sv = Saver(m, optimizer=opt)
    prefix = sv.save(str(tmp_path / "model.ckpt"))
    assert latest_checkpoint(str(tmp_path)) == prefix and os.path.exists(prefix + ".index")
    m2 = Tiny("ck1", filter_option=dr.CounterFilter(2))
    opt2 = dr.optim.AdamOptimizer(m2, lr=0.01, global_step=GlobalStep())
    step = Saver(m2, optimizer=opt2).restore(prefix)
    assert step == 6 and int(opt2.global_step) == 6 and abs(opt2.beta1_power - opt.beta1_power) < 1e-9
    probe = torch.arange(0, 45)
    assert torch.equal(m.ev.table.lookup(probe), m2.ev.table.lookup(probe))
    for s in ("m", "v"):
        assert torch.equal(m.ev.slot_values(probe, s), m2.ev.slot_values(probe, s))
    assert torch.equal(m.ev.get_frequency(probe), m2.ev.get_frequency(probe))          # filtered keys keep their counts
    assert torch.equal(m.ev.get_version(probe), m2.ev.get_version(probe))
    assert m.ev.total_count() == m2.ev.total_count() and m.ev.table.total_keys() == m2.ev.table.total_keys()
    # training continues identically
    _train(m, opt, 3); _train(m2, opt2, 3)
    assert torch.allclose(m.ev.table.lookup(probe), m2.ev.table.lookup(probe), atol=1e-6)
    assert torch.allclose(m.fc.weight, m2.fc.weight, atol=1e-6)


def test_reshard_2_to_3_partitions(tmp_path):
    m = Tiny("ck2")
    opt = dr.optim.AdagradOptimizer(m, lr=0.1, global_step=GlobalStep())
    _train(m, opt, 5, 0, 3000)
    prefix = Saver(m, optimizer=opt).save(str(tmp_path / "m.ckpt"))
    total, probe = 0, torch.arange(0, 3000)
    for p in range(3):
        mp = Tiny("ck2")
        op = dr.optim.AdagradOptimizer(mp, lr=0.1, global_step=GlobalStep())
        Saver(mp, optimizer=op, partition_id=p, partition_num=3).restore(prefix)
        keys = mp.ev.export()[0]
        assert torch.all(keys % 1000 % 3 == p)
        total += keys.numel()
        assert torch.equal(mp.ev.table.lookup(keys), m.ev.table.lookup(keys))
    assert total == m.ev.total_count()


def test_eviction_happens_at_save(tmp_path):
    m = Tiny("ck3", evict_option=dr.GlobalStepEvict(steps_to_live=2))
    opt = dr.optim.AdagradOptimizer(m, lr=0.1, global_step=GlobalStep())
    _train(m, opt, 1, 0, 20)
    n0 = m.ev.total_count()
    _train(m, opt, 5, 100, 120)
    assert m.ev.total_count() > n0
    Saver(m, optimizer=opt).save(str(tmp_path / "e.ckpt"))
    keys = m.ev.export()[0]
    assert torch.all(keys >= 100)          # everything from the first phase aged out


def test_incremental_chain_recover(tmp_path):
    m = Tiny("ck4")
    opt = dr.optim.AdagradOptimizer(m, lr=0.1, global_step=GlobalStep())
    sv = IncrementalSaver(m, optimizer=opt)
    _train(m, opt, 4, 0, 50)
    sv.save(str(tmp_path / "i.ckpt"))
    _train(m, opt, 2, 40, 70)
    p1 = sv.incremental_save(str(tmp_path / "i.ckpt"))
    _train(m, opt, 2, 60, 90)
    p2 = sv.incremental_save(str(tmp_path / "i.ckpt"))
    assert p1 != p2
    from deeprec_b200.checkpoint import BundleReader
    r = BundleReader(p2)
    incr_keys = r.read("ck4/emb-sparse_incr_keys")
    assert incr_keys.numel() > 0 and int(incr_keys.min()) >= 60          # only rows touched since the previous save
    m2 = Tiny("ck4")
    opt2 = dr.optim.AdagradOptimizer(m2, lr=0.1, global_step=GlobalStep())
    step = IncrementalSaver(m2, optimizer=opt2).recover_incr_checkpoints(str(tmp_path))
    assert step == 8
    probe = torch.arange(0, 95)
    assert torch.equal(m.ev.table.lookup(probe), m2.ev.table.lookup(probe))
    assert torch.equal(m.ev.slot_values(probe, "accumulator"), m2.ev.slot_values(probe, "accumulator"))
    assert torch.allclose(m.fc.weight, m2.fc.weight)


def test_checkpoint_option_warm_start_and_init_data_source(tmp_path):
    """CheckpointOption: a new variable (different name, different optimizer) starts from another checkpoint's tensor; init_data_source
    loads an external key/value table."""
    import deeprec_b200 as dr
    from deeprec_b200.checkpoint import Saver
    from deeprec_b200.optim import GlobalStep
    src = dr.get_embedding_variable("ws_src", 8, seed=4)
    opt = dr.optim.AdagradOptimizer([], [src], lr=0.1, global_step=GlobalStep())
    ids = torch.arange(0, 300, 3)
    for _ in range(2):
        opt.zero_grad(); (src.lookup(ids) ** 2).sum().backward(); opt.step()
    prefix = Saver(embedding_variables=[src], optimizer=opt).save(str(tmp_path / "m.ckpt"), 2)
    co = dr.CheckpointOption(ckpt_to_load_from=prefix, tensor_name_in_ckpt="ws_src")
    dst = dr.get_embedding_variable("ws_dst", 8, ev_option=dr.EmbeddingVariableOption(ckpt=co), seed=99)
    dr.optim.AdamOptimizer([], [dst], lr=0.01, global_step=GlobalStep())       # slots are created fresh, rows come from the checkpoint
    assert dst.total_count() == src.total_count() == ids.numel()
    assert torch.allclose(dst.lookup(ids).detach(), src.lookup(ids).detach()) and torch.equal(dst.get_frequency(ids), src.get_frequency(ids))
    # directory form resolves the latest checkpoint; unknown tensor names fail loudly
    dst2 = dr.get_embedding_variable("ws_src", 8, ev_option=dr.EmbeddingVariableOption(ckpt=dr.CheckpointOption(ckpt_to_load_from=str(tmp_path))), seed=5)
    assert torch.allclose(dst2.lookup(ids).detach(), src.lookup(ids).detach())
    bad = dr.get_embedding_variable("ws_bad", 8, ev_option=dr.EmbeddingVariableOption(ckpt=dr.CheckpointOption(ckpt_to_load_from=prefix, tensor_name_in_ckpt="nope")))
    with pytest.raises(KeyError):
        bad.lookup(ids)
    # external source
    table = {"keys": torch.tensor([5, 6, 7]), "values": torch.arange(24, dtype=torch.float32).view(3, 8)}
    torch.save(table, tmp_path / "ext.pt")
    ext = dr.get_embedding_variable("ws_ext", 8, ev_option=dr.EmbeddingVariableOption(ckpt=dr.CheckpointOption(init_data_source=str(tmp_path / "ext.pt"))))
    assert torch.equal(ext.lookup(torch.tensor([6])).detach(), table["values"][1:2]) and ext.total_count() == 3


def test_bundle_crc_is_ieee_crc32_also_on_the_parallel_path(tmp_path):
    """The per-tensor checksum is the standard CRC-32 (zlib's): slicing-by-8 for small tensors, parallel chunks + combine for large ones
    (> 8 MiB) -- both must equal zlib.crc32 of the bytes, and a flipped bit must be detected on read."""
    import zlib
    from deeprec_b200.checkpoint.saver import BundleReader, BundleWriter
    g = torch.Generator().manual_seed(0)
    tensors = {"tiny": torch.randn(3, generator=g), "odd": torch.randint(0, 255, (12345,), dtype=torch.uint8, generator=g),
               "large": torch.randn(5_000_011, generator=g)}                     # 20 MB, not a multiple of 8 per chunk
    prefix = str(tmp_path / "b")
    w = BundleWriter(prefix)
    for k, v in tensors.items():
        w.add(k, v)
    w.close()
    crcs = {}
    for line in open(prefix + ".index").read().splitlines()[1:]:
        tok = line.split("\t")
        crcs[tok[0]] = int(tok[-1])
    for k, v in tensors.items():
        assert crcs[k] == zlib.crc32(v.numpy().tobytes()), k
    r = BundleReader(prefix)
    assert torch.equal(r.read("large"), tensors["large"])
    r.close()
    with open(prefix + ".data", "r+b") as f:                                      # corrupt one byte in the middle of the large tensor
        f.seek(os.path.getsize(prefix + ".data") - 10_000_000); b = f.read(1); f.seek(-1, 1); f.write(bytes([b[0] ^ 0x10]))
    r = BundleReader(prefix)
    with pytest.raises(IOError):
        r.read("large")
    assert torch.equal(r.read("tiny"), tensors["tiny"])


def test_checkpoint_retention_is_per_step_across_shards(tmp_path):
    """Eight shards writing `<prefix>.ps{i}-<step>` into one directory with max_to_keep=5: every shard of the newest steps survives, whole
    steps are evicted together, and each shard finds its own latest checkpoint."""
    import os
    import torch
    from deeprec_b200.checkpoint.saver import Saver, latest_checkpoint
    d = str(tmp_path)
    savers = [Saver(extra_state={"x": torch.full((4,), float(i))}, max_to_keep=5) for i in range(8)]
    for step in range(1, 8):
        for i, sv in enumerate(savers):
            sv.save(os.path.join(d, f"model.ps{i}"), step)
    files = sorted(f for f in os.listdir(d) if f.endswith(".index"))
    assert len(files) == 8 * 5, files                                      # steps 3..7, all 8 shards each
    for step in range(3, 8):
        for i in range(8):
            assert f"model.ps{i}-{step}.index" in files
    for i in range(8):
        assert latest_checkpoint(d, base=f"model.ps{i}") == os.path.join(d, f"model.ps{i}-7")
    sv = Saver(extra_state={"x": torch.zeros(4)})
    sv.restore(latest_checkpoint(d, base="model.ps5"))
    assert float(sv.extra["x"][0]) == 5.0


def test_sp_owner_matches_the_device_hash():
    """checkpoint/engine_ckpt.sp_owner (torch int64 arithmetic) == csrc/cuda/sparse_pipeline.cu::sp_owner (uint64 arithmetic)."""
    import torch
    from deeprec_b200.checkpoint.engine_ckpt import sp_owner
    M = (1 << 64) - 1

    def mix64(x):
        x ^= x >> 30; x = (x * 0xbf58476d1ce4e5b9) & M
        x ^= x >> 27; x = (x * 0x94d049bb133111eb) & M
        x ^= x >> 31
        return x
    g = torch.Generator().manual_seed(0)
    keys = torch.cat([torch.randint(-(1 << 62), 1 << 62, (500,), generator=g), torch.arange(-5, 50), torch.tensor([(1 << 63) - 1, -(1 << 63) + 2])])
    for W in (2, 3, 8):
        want = [(mix64(((int(k) & M) ^ 0x5bd1e9955bd1e995)) >> 33) % W for k in keys.tolist()]
        assert sp_owner(keys, W).tolist() == want
    assert sp_owner(keys, 1).sum().item() == 0
