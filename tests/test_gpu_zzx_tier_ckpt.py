"""Engine checkpoints of multi-tier tables (checkpoint/engine_ckpt.py): full and incremental saves carry BOTH tiers (rows demoted to the DRAM
tier keep their dirty bit), restore puts each row back into its tier, and a tiered checkpoint restores into an untiered engine.
Written after the round's GPU budget was spent: sorts late on purpose.  Reference: hbm_dram_storage.h Save() walks both tiers."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _engine(tiered):
    import deeprec_b200 as dr
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    return criteo_engine(build_model("deepfm", device="cuda"), 512, table_rows=[50, 6000, 7, 300] + [97] * 22, learning_rate=0.05, tiered=tiered)


def _batches(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cards, B, out = [50, 6000, 7, 300] + [97] * 22, 512, []
    for s in range(n):
        lo = (s % 4) * 1500                                    # table 1's working set rotates: 6000 distinct ids >> 1024 cache rows
        ids = torch.stack([torch.randint(0, c, (B,), device="cuda", generator=g) for c in cards])
        ids[1] = torch.randint(lo, lo + 1500, (B,), device="cuda", generator=g)
        out.append((ids, (torch.rand(B, device="cuda", generator=g) < 0.3).float(), {"dense": torch.rand(B, 13, device="cuda", generator=g)}))
    return out


def _run(eng, batches):
    if eng.tiers:
        eng.prefetch(batches[0][0])
    for s, (ids, y, dense) in enumerate(batches):
        eng.load_batch(ids, y, dense); eng.train_step()
        if eng.tiers and s + 1 < len(batches):
            eng.prefetch(batches[s + 1][0])
    torch.cuda.synchronize()


def test_full_and_incremental_checkpoints_carry_both_tiers(tmp_path):
    tiered = {1: {"cache_rows": 1024, "strategy": 0}}
    a = _engine(tiered)
    _run(a, _batches(12, 1))
    mgr = a.tiers[1][0]
    assert mgr.stats()["demoted_rows"] > 0                      # part of table 1 lives in the DRAM tier now
    path = str(tmp_path / "eng")
    a.save(path)
    _run(a, _batches(6, 2))                                     # more steps: rows change in both tiers, new demotions
    a.save(path, incremental=True)
    probe = torch.arange(0, 6000, 5, device="cuda")
    want = mgr.lookup(probe)
    small = torch.arange(0, 300, device="cuda")

    b = _engine(tiered)
    step = b.restore(path)
    assert step > 0
    got = b.tiers[1][0].lookup(probe)
    assert torch.allclose(got, want, atol=1e-6), (got - want).abs().max()
    assert torch.allclose(b.tables[3].lookup(small), a.tables[3].lookup(small), atol=1e-6)
    assert b.tiers[1][0].host.size() > 0                        # the DRAM tier was restored as a DRAM tier

    c = _engine(None)                                           # the same checkpoint into an engine that keeps everything in HBM
    c.restore(path)
    assert torch.allclose(c.tables[1].lookup(probe), want, atol=1e-6)
    # and training continues identically from the restored state on both
    nxt = _batches(3, 3)
    _run(a, nxt); _run(b, nxt)
    assert abs(a.loss_value() - b.loss_value()) < 1e-4
