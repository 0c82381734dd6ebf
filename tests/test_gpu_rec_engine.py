"""GPU: FusedRecEngine (unique-first sparse pipeline + autograd dense net + fused optimizers, one CUDA graph) against a plain PyTorch
fp32 replica of the same model with nn.Embedding tables (DeepFM: BASELINE config #3; DIN: config #4)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _adagrad(params, lr):
    return torch.optim.Adagrad(params, lr=lr, initial_accumulator_value=0.1, eps=0.0)


@pytest.mark.parametrize("graph", [False, True])
def test_deepfm_engine_matches_torch_replica(graph):
    import deeprec_b200 as dr
    from deeprec_b200.models.rec_engine import criteo_engine
    from deeprec_b200.models.zoo import build_model
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    B, cards = 1024, [50, 1000, 7, 300] + [97] * 22
    model = build_model("deepfm", device="cuda")
    eng = criteo_engine(model, B, table_rows=cards, learning_rate=0.05)
    ref_net = copy.deepcopy(eng.net.inner)                    # dense part (parameters are independent copies)
    tables = torch.nn.ModuleList([torch.nn.Embedding(c, 16, device="cuda") for c in cards])
    with torch.no_grad():
        for t, emb in enumerate(tables):
            dm = eng.tables[t].default_matrix
            emb.weight.copy_(dm[torch.arange(cards[t], device="cuda") % dm.shape[0]])
    opt = _adagrad(list(ref_net.parameters()) + list(tables.parameters()), 0.05)
    le, lr_ = [], []
    steps = 6
    for s in range(steps):
        dense = torch.rand(B, 13, device="cuda") * 3
        ids = torch.stack([torch.randint(0, c, (B,), device="cuda") for c in cards])
        y = (torch.rand(B, device="cuda") < 0.3).float()
        eng.load_batch(ids, y, {"dense": dense})
        if graph and s == 0:
            eng.capture(warmup=1)                             # 1 eager step on this batch, then capture (the capture itself does not run)
        else:
            eng.train_step()
        le.append(eng.loss_value())
        embs = torch.stack([tables[t](ids[t]) for t in range(26)], 1)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(ref_net.logits(dense, embs).float(), y)
        opt.zero_grad(); loss.backward(); opt.step()
        lr_.append(loss.item())
    for a, b in zip(le, lr_):
        assert abs(a - b) < 0.03 * max(1.0, abs(b)), (le, lr_)
    keys = ids[1][:64]
    assert (eng.tables[1].lookup(keys) - tables[1].weight[keys]).abs().max().item() < 0.05
    assert eng.tables[0].overflowed() == 0


def test_din_engine_trains_with_padding_and_admission():
    import deeprec_b200 as dr
    from deeprec_b200.data import taobao_batch
    from deeprec_b200.models.rec_engine import din_engine, din_ids
    from deeprec_b200.models.zoo import build_model
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    B, L = 512, 50
    model = build_model("din", device="cuda")
    eng = din_engine(model, B, L, table_rows=(5000, 20000, 100), learning_rate=0.05, filter_freq=2)
    losses = []
    b = taobao_batch(B, L, 5000, 20000, 100, seed=0)                  # one batch, repeated: the model must fit it
    ids, y = din_ids(b).cuda(), b["labels"].cuda()
    for s in range(40):
        eng.load_batch(ids, y)
        if s == 2:
            eng.capture(warmup=1)
        else:
            eng.train_step()
        losses.append(eng.loss_value())
    assert all(l == l for l in losses) and losses[-1] < losses[0] - 0.02, losses
    # padding positions never create keys; un-admitted keys (seen once) own no row
    item = eng.tables[1]
    assert item.size() > 0 and item.total_keys() >= item.size()
    p = eng.predict()
    assert p.shape == (B,) and bool(((p >= 0) & (p <= 1)).all())


def test_engine_checkpoint_full_and_incremental_roundtrip(tmp_path):
    """DLRMEngine.save / restore: full checkpoint + delta chain reproduce the live engine (dense block, table rows incl. optimizer slots,
    frequencies, global step), and training continues identically from the restored state."""
    from deeprec_b200.models.dlrm_engine import DLRMConfig, DLRMEngine
    cards = [50, 1000, 7, 300] + [97] * 22
    cfg = DLRMConfig(batch_size=512, cardinalities=cards, learning_rate=0.05)
    torch.manual_seed(1)
    batches = [(torch.rand(512, 13, device="cuda"), torch.stack([torch.randint(0, c, (512,), device="cuda") for c in cards]),
                (torch.rand(512, device="cuda") < 0.3).float()) for _ in range(7)]
    a = DLRMEngine(cfg)
    for b in batches[:3]:
        a.load_batch(*b); a.train_step()
    path = str(tmp_path / "dlrm")
    a.save(path)
    for b in batches[3:5]:
        a.load_batch(*b); a.train_step()
    a.save(path, incremental=True)                      # only rows touched by steps 4-5 + the dense block
    b2 = DLRMEngine(cfg)
    step = b2.restore(path)
    assert step == 5 and b2.global_step() == 5
    assert torch.equal(a.params, b2.params) and torch.equal(a.s0, b2.s0)
    probe = torch.arange(0, 1000, device="cuda")
    for t in (0, 1, 3):
        assert torch.equal(a.tables[t].get_freq(probe), b2.tables[t].get_freq(probe))
        assert torch.allclose(a.tables[t].lookup(probe), b2.tables[t].lookup(probe))
        assert torch.allclose(a.tables[t].lookup_slot(probe, 1), b2.tables[t].lookup_slot(probe, 1))
    la, lb = [], []
    for bt in batches[5:]:
        for e, l in ((a, la), (b2, lb)):
            e.load_batch(*bt); e.train_step(); l.append(e.loss_value())
    assert max(abs(x - y) for x, y in zip(la, lb)) < 2e-3, (la, lb)
    # a full-only restore lands on the full checkpoint's step
    c = DLRMEngine(cfg)
    assert c.restore(path, replay_incremental=False) == 3
