"""Feature columns + lookup API (embedding_variable_ops_test.py: shared / weighted / sequence / adaptive / multihash columns)."""
import torch

import deeprec_b200 as dr
from deeprec_b200 import feature_column as fc
from deeprec_b200.optim import GlobalStep


def test_embedding_lookup_sparse_combiners_match_manual():
    ev = dr.get_embedding_variable("fc_comb", 4, seed=1)
    vals = torch.tensor([1, 2, 3, 4, 5]); off = torch.tensor([0, 2, 2, 5])
    sp = dr.SparseIds.from_offsets(vals, off)
    rows = ev.lookup(vals).detach()
    s = dr.embedding_lookup_sparse(ev, sp, combiner="sum")
    assert torch.allclose(s[0], rows[0] + rows[1]) and torch.all(s[1] == 0) and torch.allclose(s[2], rows[2:].sum(0))
    m = dr.embedding_lookup_sparse(ev, sp, combiner="mean")
    assert torch.allclose(m[2], rows[2:].mean(0)) and torch.all(m[1] == 0)
    w = torch.tensor([1.0, 3.0, 1.0, 1.0, 2.0])
    q = dr.embedding_lookup_sparse(ev, sp, w, combiner="sqrtn")
    assert torch.allclose(q[0], (rows[0] + 3 * rows[1]) / (10 ** 0.5), atol=1e-6)
    safe = dr.safe_embedding_lookup_sparse(ev, dr.SparseIds(torch.tensor([7, -1, 9]), torch.tensor([0, 0, 2]), 3), combiner="sum", default_id=0)
    assert torch.allclose(safe[1], ev.lookup(torch.tensor([0])).detach()[0])


def test_input_layer_shared_weighted_sequence_numeric_multihash_adaptive():
    cols = []
    a = fc.categorical_column_with_embedding("user", ev_option=dr.EmbeddingVariableOption(filter_option=dr.CounterFilter(1)))
    b = fc.categorical_column_with_embedding("item")
    cols += fc.shared_embedding_columns([fc.categorical_column_with_embedding("click"), fc.categorical_column_with_embedding("buy")], 4, combiner="sum")
    cols.append(fc.embedding_column(a, 4))
    cols.append(fc.embedding_column(fc.weighted_categorical_column(b, "item_w"), 4, combiner="sum"))
    cols.append(fc.embedding_column(fc.sequence_categorical_column_with_embedding("hist"), 4))
    cols.append(fc.embedding_column(fc.categorical_column_with_multihash("mh", [[10, 4], [7, 4]]), 4))
    cols.append(fc.embedding_column(fc.categorical_column_with_adaptive_embedding("ad", 100), 4))
    cols.append(fc.embedding_column(fc.categorical_column_with_hash_bucket("hb", 50), 4))
    cols.append(fc.numeric_column("age"))
    layer = fc.InputLayer(cols, name="il")
    B = 3
    feats = {"click": torch.tensor([[1, 2], [3, -1], [-1, -1]]), "buy": torch.tensor([[1], [2], [3]]), "user": torch.tensor([5, 6, 7]),
             "item": torch.tensor([[1, 2], [3, -1], [4, 5]]), "item_w": torch.tensor([[0.5, 2.0], [1.0, 0.0], [1.0, 1.0]]),
             "hist": torch.tensor([[1, 2, -1], [3, -1, -1], [4, 5, 6]]), "mh": torch.tensor([3, 69, 12]), "ad": torch.tensor([11, 12, 13]),
             "hb": torch.tensor([123456789, 5, 6]), "age": torch.tensor([[0.1], [0.2], [0.3]])}
    outs = {}
    x = layer(feats, adaptive_mask_tensors={"ad": torch.tensor([True, False, True])}, cols_to_output_tensors=outs)
    assert x.shape == (B, 4 + 4 + 4 + 4 + 3 * 4 + 4 + 4 + 4 + 1)
    # shared table: click and buy columns resolve to the same EmbeddingVariable
    assert len({id(layer.tables[layer._key_of[id(c)]]) for c in cols[:2]}) == 1
    assert torch.all(outs["hist_embedding"][0, 2] == 0)                      # padded step
    opt = dr.optim.AdagradOptimizer(layer, lr=0.1, global_step=GlobalStep())
    x.sum().backward(); opt.step()
    assert all(e.total_count() > 0 for e in layer.embedding_variables())


def test_group_embedding_column_scope_matches_ungrouped():
    def build(grouped):
        cs = []
        ctx = fc.group_embedding_column_scope("g") if grouped else None
        if ctx:
            ctx.__enter__()
        for k in ("f1", "f2", "f3"):
            cs.append(fc.embedding_column(fc.categorical_column_with_embedding(k), 8, combiner="mean"))
        if ctx:
            ctx.__exit__(None, None, None)
        return fc.InputLayer(cs, name="grp")
    feats = {"f1": torch.tensor([[1, 2], [3, -1]]), "f2": torch.tensor([[4, -1], [5, 6]]), "f3": torch.tensor([7, 8])}
    a, b = build(True)(feats), build(False)(feats)
    assert torch.allclose(a, b)
