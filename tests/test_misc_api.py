"""Smaller API pieces: CSV helper ops, scatter_add, sample-aware compression, streaming AUC, tracing."""
import torch

import deeprec_b200 as dr
from deeprec_b200.data.csv_ops import (sparse_valid_cutoff, string_split_and_pad, string_split_and_pad_ids, string_to_hash_id, trans_csv_id2dense,
                                        trans_csv_id2sparse, trans_csv_kv2dense, trans_csv_kv2sparse, trans_csv_to_dense)
from deeprec_b200.serving.sample_aware import enable_sample_awared_graph_compression
from deeprec_b200.utils import StreamingAUC, Timeline


def test_csv_ops():
    assert string_split_and_pad(["a,b,c", "d", ""], 2, default_value="<pad>") == [["a", "b"], ["d", "<pad>"], ["<pad>", "<pad>"]]
    # the examples of kernels/trans_csv_ali_ops.cc:8-55 (the second dimension of every result is max_id; ids / keys live in [0, max_id))
    sp = trans_csv_id2sparse(["2,10", "7", "0,8"], max_id=12)
    assert sp.values.tolist() == [2, 10, 7, 0, 8] and sp.row_ids.tolist() == [0, 0, 1, 2, 2] and sp.batch_size == 3 and sp.dense_shape == (3, 12)
    sp = trans_csv_id2sparse(["1, 2,,9", "", "3"], max_id=None)                      # empty tokens / records skipped, max_id detected
    assert sp.values.tolist() == [1, 2, 9, 3] and sp.row_ids.tolist() == [0, 0, 0, 2] and sp.dense_shape == (3, 10)
    assert trans_csv_id2sparse(["4"], max_id=5, id_as_value=False, default_value=2.0).weights.tolist() == [2.0]
    assert trans_csv_id2dense(["2,1", "3", "0,2"], max_id=4, default_value=1).tolist() == [[0, 1, 1, 0], [0, 0, 0, 1], [1, 0, 1, 0]]
    kv = trans_csv_kv2sparse(["2:2.0,10:0.1", "7:-0.7", "8:0.8"], max_id=12)
    assert kv.values.tolist() == [2, 10, 7, 8] and kv.row_ids.tolist() == [0, 0, 1, 2] and torch.allclose(kv.weights, torch.tensor([2.0, 0.1, -0.7, 0.8]))
    d = trans_csv_kv2dense(["2:0.2,1:0.1", "3:-0.3", "0:0.4,2:0.2"], max_id=4)
    assert torch.allclose(d, torch.tensor([[0.0, 0.1, 0.2, 0.0], [0.0, 0.0, 0.0, -0.3], [0.4, 0.0, 0.2, 0.0]]))
    assert torch.allclose(trans_csv_to_dense(["0.2,0.1", "-0.3", "0.4,0.2"], max_id=4), torch.tensor([[0.2, 0.1, 0, 0], [-0.3, 0, 0, 0], [0.4, 0.2, 0, 0]]))
    assert trans_csv_to_dense(["1;2;3", "4"], field_delim=";").shape == (2, 3)
    assert string_split_and_pad_ids(["5,6,7,8", "", "9"], 3).tolist() == [[5, 6, 7], [-1, -1, -1], [9, -1, -1]]
    for bad in (lambda: trans_csv_id2sparse(["1,99"], max_id=10), lambda: trans_csv_id2sparse(["1,x"]), lambda: trans_csv_kv2dense(["1=2"], 4),
                lambda: trans_csv_to_dense(["1,abc"])):
        try:
            bad(); raise AssertionError("malformed / out-of-range input must raise")
        except ValueError:
            pass
    big = [",".join(str((i * 7 + j) % 1000) for j in range(i % 9)) for i in range(5000)]         # a real batch: parallel count + fill passes
    sp = trans_csv_id2sparse(big, max_id=1000)
    assert sp.values.numel() == sum(i % 9 for i in range(5000)) and torch.equal(torch.bincount(sp.row_ids, minlength=5000), torch.tensor([i % 9 for i in range(5000)]))
    sp = dr.SparseIds(torch.arange(7), torch.tensor([0, 0, 0, 0, 1, 1, 2]), 3)
    assert sparse_valid_cutoff(sp, 2, "left").values.tolist() == [0, 1, 4, 5, 6]
    assert sparse_valid_cutoff(sp, 2, "right").values.tolist() == [2, 3, 4, 5, 6]
    assert string_to_hash_id("abc") == string_to_hash_id("abc") != string_to_hash_id("abd")


def test_scatter_add_creates_and_accumulates():
    ev = dr.get_embedding_variable("sa", 4, seed=1)
    base = ev.lookup(torch.tensor([7])).detach().clone()
    ev.scatter_add(torch.tensor([7, 7]), torch.ones(2, 4))
    assert torch.allclose(ev.lookup(torch.tensor([7])).detach(), base + 2.0) and ev.total_count() == 1


def test_sample_aware_compression_matches_tiled():
    torch.manual_seed(0)
    un, it, hd = torch.nn.Linear(6, 4), torch.nn.Linear(5, 4), torch.nn.Linear(8, 1)
    sac = enable_sample_awared_graph_compression(un, it, hd)
    u, items = torch.randn(1, 6), torch.randn(9, 5)
    ref = hd(torch.cat([un(u.expand(9, -1)), it(items)], -1)).squeeze(-1)
    assert torch.allclose(sac(u, items), ref, atol=1e-6)


def test_auc_and_timeline(tmp_path):
    auc = StreamingAUC()
    y = (torch.rand(5000) < 0.3).float()
    auc.update(torch.where(y > 0, torch.rand(5000) * 0.5 + 0.5, torch.rand(5000) * 0.5), y)
    assert auc.result() > 0.99
    auc2 = StreamingAUC(); auc2.update(torch.rand(5000), y)
    assert 0.45 < auc2.result() < 0.55
    tl = Timeline()
    with tl.span("step"):
        sum(range(1000))
    tl.save(str(tmp_path / "t.json"))
    import json
    assert json.load(open(tmp_path / "t.json"))["traceEvents"][0]["name"] == "step"


import pytest


@pytest.mark.parametrize("name", ["gradientdescent", "adagrad", "adam", "adamw", "adamasync", "adagraddecay", "ftrl"])
def test_multi_tensor_dense_updates_equal_the_per_parameter_rules(name):
    """The ``torch._foreach_*`` dense paths must reproduce the scalar rules (which the sparse kernels are tested against)."""
    import deeprec_b200 as dr
    from deeprec_b200.optim import GlobalStep, make_optimizer
    from deeprec_b200.optim.optimizers import DeepRecOptimizer
    torch.manual_seed(0)

    def net():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3, bias=False), torch.nn.ReLU(), torch.nn.Linear(3, 1))

    a, b = net(), net()
    kw = dict(lr=0.05)
    if name == "adagraddecay":
        kw.update(accumulator_decay_step=2, accumulator_decay_rate=0.5)
    oa = make_optimizer(name, a, None, global_step=GlobalStep(), **kw)
    ob = make_optimizer(name, b, None, global_step=GlobalStep(), **kw)
    ob._dense_update_many = lambda ps, gs, sts, group, hp: DeepRecOptimizer._dense_update_many(ob, ps, gs, sts, group, hp)   # scalar path
    for step in range(4):
        x = torch.randn(16, 5)
        if step == 2:
            a[4].weight.requires_grad_(False); b[4].weight.requires_grad_(False)      # a parameter without gradient is skipped by both
        for m, o in ((a, oa), (b, ob)):
            o.zero_grad(); m(x).pow(2).mean().backward(); o.step()
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa, pb, atol=1e-6, rtol=1e-5), (name, (pa - pb).abs().max())


def test_summary_hook_writes_tensorboard_events(tmp_path):
    import glob
    import deeprec_b200 as dr
    from deeprec_b200.utils import SummaryHook, Trainer
    dr.embedding_variable.clear_registry()
    ev = dr.get_embedding_variable("sum/emb", 4, seed=1)
    head = torch.nn.Linear(4, 1)
    model = torch.nn.ModuleList([ev, head])
    opt = dr.optim.AdagradOptimizer(model, lr=0.1, global_step=dr.optim.GlobalStep())
    tr = Trainer(model, opt, lambda m, ids: m[1](m[0].lookup(ids)).pow(2).mean(), log_every_n_steps=0, hooks=[SummaryHook(str(tmp_path), every_n_steps=2)])
    tr.fit((torch.randint(0, 50, (16,)) for _ in range(6)))
    files = glob.glob(str(tmp_path / "events.out.tfevents.*"))
    assert files
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    acc = EventAccumulator(str(tmp_path)); acc.Reload()
    tags = acc.Tags()["scalars"]
    assert "loss" in tags and "embedding_variable/sum/emb/rows" in tags and "global_step/sec" in tags
    assert [e.step for e in acc.Scalars("loss")] == [2, 4, 6] and acc.Scalars("embedding_variable/sum/emb/rows")[-1].value == ev.total_count()


@pytest.mark.parametrize("name", ["adagrad", "adam", "adamasync", "adagraddecay", "ftrl", "gradientdescent"])
def test_sparse_gradients_of_plain_embeddings_are_applied_lazily(name):
    """nn.Embedding(sparse=True) under the DeepRec optimizers = SparseApply* on a non-EV variable: when every row is touched the result
    equals the dense update; rows that are not touched keep their values AND their slots (no decay)."""
    from deeprec_b200.optim import GlobalStep, make_optimizer
    torch.manual_seed(0)
    kw = dict(lr=0.05)
    if name == "adagraddecay":
        kw.update(accumulator_decay_step=2, accumulator_decay_rate=0.5)
    es, ed = torch.nn.Embedding(12, 4, sparse=True), torch.nn.Embedding(12, 4)
    ed.weight.data.copy_(es.weight.data)
    os_, od = make_optimizer(name, es, None, global_step=GlobalStep(), **kw), make_optimizer(name, ed, None, global_step=GlobalStep(), **kw)
    allids = torch.arange(12)
    for step in range(3):                                   # all rows touched: sparse == dense
        t = torch.randn(12, 4)
        for e, o in ((es, os_), (ed, od)):
            o.zero_grad(); ((e(allids) - t) ** 2).sum().backward(); o.step()
    assert torch.allclose(es.weight, ed.weight, atol=1e-6)
    before = es.weight.detach().clone()
    st_before = {k: v.clone() for k, v in os_.state[es.weight].items() if torch.is_tensor(v)}
    os_.zero_grad(); (es(torch.tensor([2, 2, 5])) ** 2).sum().backward(); os_.step()
    touched = torch.zeros(12, dtype=torch.bool); touched[[2, 5]] = True
    assert torch.equal(es.weight.detach()[~touched], before[~touched]) and not torch.allclose(es.weight.detach()[touched], before[touched])
    for k, v in st_before.items():
        assert torch.equal(os_.state[es.weight][k][~touched], v[~touched]), k
