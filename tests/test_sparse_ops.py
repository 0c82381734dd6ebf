"""ops/sparse_ops.py on CPU tensors (the torch expressions that are also the oracle of the CUDA kernels, tests/test_gpu_zzy_sparse_utils.py)
against brute-force python loops.  Reference semantics: tf.sparse.retain + fill_empty_rows in safe_embedding_lookup_sparse
(python/ops/embedding_ops.py:838), tf.sparse.slice / reshape, tf.sparse.segment_{sum,mean,sqrt_n}."""
import math

import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.ops.sparse_ops import (sparse_fill_empty_rows, sparse_prune_fill, sparse_reshape, sparse_segment_mean, sparse_segment_sqrt_n,
                                         sparse_segment_sum, sparse_slice)


def _random_sp(B, L, seed, weights=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(-2, 50, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    mask = torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)
    rows = torch.arange(B).unsqueeze(1).expand(B, L)
    w = (torch.rand(B, L, generator=g) - 0.2) if weights else None
    return dr.SparseIds(ids[mask], rows[mask], B, w[mask] if weights else None)


@pytest.mark.parametrize("weights", [True, False])
def test_prune_fill_matches_a_python_loop(weights):
    sp = _random_sp(37, 6, 0, weights)
    out, empty = sparse_prune_fill(sp, default_id=7)
    exp = []
    for b in range(sp.batch_size):
        row = [(int(v), float(w) if weights else 1.0) for v, r, w in zip(sp.values, sp.row_ids, sp.weights if weights else [1.0] * sp.values.numel())
               if int(r) == b and int(v) >= 0 and (not weights or float(w) > 0)]
        assert bool(empty[b]) == (len(row) == 0)
        exp += [(b, v, w) for v, w in row] or [(b, 7, 1.0)]
    assert out.row_ids.tolist() == [e[0] for e in exp] and out.values.tolist() == [e[1] for e in exp]
    if weights:
        assert torch.allclose(out.weights, torch.tensor([e[2] for e in exp]))
    # no default id: empty rows stay empty; no pruning: invalid ids stay
    out2, _ = sparse_prune_fill(sp, default_id=None)
    assert out2.values.numel() == sum(1 for e in exp if not (e[1] == 7 and e[2] == 1.0 and bool(empty[e[0]])))
    out3, e3 = sparse_fill_empty_rows(sp, 3)
    assert out3.values.numel() == sp.values.numel() + int(e3.sum()) and (out3.row_ids[1:] >= out3.row_ids[:-1]).all()


def test_slice_and_reshape():
    g = torch.Generator().manual_seed(1)
    shape = [6, 5, 4]
    dense = (torch.rand(shape, generator=g) < 0.3) * torch.randint(1, 100, shape, generator=g)
    idx = dense.nonzero(); val = dense[dense != 0]
    oi, ov, oshape = sparse_slice(idx, val, shape, [1, 0, 2], [3, 4, 5])
    sub = dense[1:4, 0:4, 2:4]
    assert oshape == [3, 4, 2] and torch.equal(oi, sub.nonzero()) and torch.equal(ov, sub[sub != 0])
    ri, rshape = sparse_reshape(idx, shape, [10, -1])
    assert rshape == [10, 12] and torch.equal(ri, dense.reshape(10, 12).nonzero())
    with pytest.raises(ValueError):
        sparse_reshape(idx, shape, [7, -1])


@pytest.mark.parametrize("mode", ["sum", "mean", "sqrtn"])
def test_sparse_segment_reductions(mode):
    g = torch.Generator().manual_seed(2)
    data = torch.randn(20, 8, generator=g, requires_grad=True)
    indices = torch.randint(0, 20, (33,), generator=g)
    seg = torch.sort(torch.randint(0, 9, (33,), generator=g)).values
    fn = {"sum": sparse_segment_sum, "mean": sparse_segment_mean, "sqrtn": sparse_segment_sqrt_n}[mode]
    out = fn(data, indices, seg, 10)
    for s in range(10):
        sel = [int(i) for i, t in zip(indices, seg) if int(t) == s]
        ref = data.detach()[sel].sum(0) if sel else torch.zeros(8)
        if sel and mode != "sum":
            ref = ref / (len(sel) if mode == "mean" else math.sqrt(len(sel)))
        assert torch.allclose(out[s], ref, atol=1e-5)
    out.sum().backward()
    assert data.grad is not None and data.grad.abs().sum() > 0
