"""csrc/cuda/sparse_utils.cu (prune + fill-empty-rows, COO slice / reshape, sparse segment reductions) against the CPU torch expressions of
ops/sparse_ops.py.  Written after the round's GPU budget was spent: this file sorts last on purpose (SIMT + cub kernels, no barrier protocols)."""
import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200.ops.sparse_ops import (sparse_fill_empty_rows, sparse_prune_fill, sparse_reshape, sparse_segment_mean, sparse_segment_sqrt_n,
                                         sparse_segment_sum, sparse_slice)

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


def _random_sp(B, L, seed, weights=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(-2, 5000, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    mask = torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)
    rows = torch.arange(B).unsqueeze(1).expand(B, L)
    w = (torch.rand(B, L, generator=g) - 0.2) if weights else None
    return dr.SparseIds(ids[mask], rows[mask], B, w[mask] if weights else None)


@pytest.mark.parametrize("B,L,weights", [(37, 6, True), (4096, 9, True), (20000, 3, False), (5, 1, True)])
def test_prune_fill_cuda_matches_cpu(B, L, weights):
    sp = _random_sp(B, L, B + L, weights)
    for default_id, prune in ((7, True), (None, True), (3, False)):
        ref, ref_empty = sparse_prune_fill(sp, default_id, prune)
        got, got_empty = sparse_prune_fill(sp.to("cuda"), default_id, prune)
        assert torch.equal(got.values.cpu(), ref.values) and torch.equal(got.row_ids.cpu(), ref.row_ids) and torch.equal(got_empty.cpu(), ref_empty)
        if weights:
            assert torch.allclose(got.weights.cpu(), ref.weights)
    e, ind = sparse_fill_empty_rows(dr.SparseIds(torch.empty(0, dtype=torch.int64, device="cuda"), torch.empty(0, dtype=torch.int64, device="cuda"), 4), 9)
    assert e.values.tolist() == [9, 9, 9, 9] and e.row_ids.tolist() == [0, 1, 2, 3] and bool(ind.all())


def test_slice_and_reshape_cuda_match_cpu():
    g = torch.Generator().manual_seed(1)
    shape = [64, 33, 17]
    dense = (torch.rand(shape, generator=g) < 0.2) * torch.randint(1, 100, shape, generator=g)
    idx = dense.nonzero(); val = dense[dense != 0]
    for v in (val, val.float()):
        ri, rv, rs = sparse_slice(idx, v, shape, [3, 0, 5], [40, 20, 100])
        gi, gv, gs = sparse_slice(idx.cuda(), v.cuda(), shape, [3, 0, 5], [40, 20, 100])
        assert gs == rs and torch.equal(gi.cpu(), ri) and torch.equal(gv.cpu(), rv)
    r1, s1 = sparse_reshape(idx, shape, [33, -1, 2])
    g1, s2 = sparse_reshape(idx.cuda(), shape, [33, -1, 2])
    assert s1 == s2 and torch.equal(g1.cpu(), r1)


@pytest.mark.parametrize("mode", ["sum", "mean", "sqrtn"])
def test_sparse_segment_reductions_cuda_match_cpu(mode):
    g = torch.Generator().manual_seed(2)
    fn = {"sum": sparse_segment_sum, "mean": sparse_segment_mean, "sqrtn": sparse_segment_sqrt_n}[mode]
    data = torch.randn(3000, 48, generator=g)
    indices = torch.randint(0, 3000, (20000,), generator=g)
    seg = torch.sort(torch.randint(0, 700, (20000,), generator=g)).values
    d_cpu = data.clone().requires_grad_(True); d_gpu = data.cuda().requires_grad_(True)
    ref = fn(d_cpu, indices, seg, 701); got = fn(d_gpu, indices.cuda(), seg.cuda(), 701)
    assert torch.allclose(got.cpu(), ref, atol=1e-4, rtol=1e-4)
    w = torch.randn(701, 48, generator=g)
    (ref * w).sum().backward(); (got * w.cuda()).sum().backward()
    assert torch.allclose(d_gpu.grad.cpu(), d_cpu.grad, atol=1e-3, rtol=1e-3)
