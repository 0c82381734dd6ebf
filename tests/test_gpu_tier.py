"""GPU: native multi-tier storage (csrc/cuda/tier_kernels.cu + ops/tier_manager.py): a tiny HBM cache over the host DRAM tier trains to
the same result as an all-HBM table; cold rows are demoted (LFU / LRU histogram threshold) and promoted back by the prefetch path."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("strategy", [0, 1])
def test_tiered_engine_matches_untiered(strategy):
    import deeprec_b200 as dr
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    B, cards = 512, [50, 6000, 7, 300] + [97] * 22
    engs = []
    for tiered in (None, {1: {"cache_rows": 1024, "strategy": strategy}}):
        dr.embedding_variable.clear_registry()
        torch.manual_seed(0)
        model = build_model("deepfm", device="cuda")
        engs.append(criteo_engine(model, B, table_rows=cards, learning_rate=0.05, tiered=tiered))
    ref, tier = engs
    mgr = tier.tiers[1][0]
    torch.manual_seed(1)
    batches = []
    for s in range(14):
        lo = (s % 4) * 1500                                    # table 1's working set rotates: 6000 distinct ids >> 1024 cache rows
        ids = torch.stack([torch.randint(0, c, (B,), device="cuda") for c in cards])
        ids[1] = torch.randint(lo, lo + 1500, (B,), device="cuda")
        batches.append((ids, (torch.rand(B, device="cuda") < 0.3).float(), {"dense": torch.rand(B, 13, device="cuda")}))
    la, lb = [], []
    tier.prefetch(batches[0][0])
    for s, (ids, y, dense) in enumerate(batches):
        for e, l in ((ref, la), (tier, lb)):
            e.load_batch(ids, y, dense)
            if e is tier and s + 1 < len(batches):
                pass
            e.train_step()
            l.append(e.loss_value())
        if s + 1 < len(batches):
            tier.prefetch(batches[s + 1][0])                   # one batch ahead, while (conceptually) the step runs
    assert max(abs(a - b) for a, b in zip(la, lb)) < 2e-3, (la, lb)
    st = mgr.stats()
    assert st["demoted_rows"] > 0 and st["promoted_rows"] > 0 and st["evict_passes"] > 0, st
    assert st["hbm_rows"] <= 1024 + 2 * B + 4096
    probe = torch.arange(0, 6000, 7, device="cuda")
    assert torch.allclose(mgr.lookup(probe), ref.tables[1].lookup(probe), atol=1e-4)
    assert tier.tables[1].overflowed() == 0
