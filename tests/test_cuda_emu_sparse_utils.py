"""csrc/cuda/sparse_utils.cu executed on the CPU through the CUDA-on-CPU emulation (csrc/cuda/emu/cuda_emu.h: one host thread per CUDA
thread, same sources, same python wrappers) against the torch expressions of ops/sparse_ops.py -- the GPU test
tests/test_gpu_zzy_sparse_utils.py with the emulation in place of the device, at sizes a CPU box can afford."""
import os

import pytest
import torch

import deeprec_b200 as dr
from deeprec_b200 import _native
from deeprec_b200.ops.sparse_ops import (sparse_fill_empty_rows, sparse_prune_fill, sparse_reshape, sparse_segment_mean, sparse_segment_sqrt_n,
                                         sparse_segment_sum, sparse_slice)

pytestmark = [pytest.mark.timeout(600)]


def _random_sp(B, L, seed, weights=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(-2, 5000, (B, L), generator=g)
    lens = torch.randint(0, L + 1, (B,), generator=g)
    mask = torch.arange(L).unsqueeze(0) < lens.unsqueeze(1)
    rows = torch.arange(B).unsqueeze(1).expand(B, L)
    w = (torch.rand(B, L, generator=g) - 0.2) if weights else None
    return dr.SparseIds(ids[mask], rows[mask], B, w[mask] if weights else None)


@pytest.mark.parametrize("B,L,weights", [(37, 6, True), (5, 1, True)] if os.environ.get("DEEPREC_EMU_QUICK") == "1" else [(37, 6, True), (700, 9, True), (1500, 3, False), (5, 1, True)])
def test_prune_fill_emulated_kernels_match_cpu(B, L, weights):
    sp = _random_sp(B, L, B + L, weights)
    for default_id, prune in ((7, True), (None, True), (3, False)):
        ref, ref_empty = sparse_prune_fill(sp, default_id, prune)
        with _native.cuda_emulation():
            got, got_empty = sparse_prune_fill(sp, default_id, prune)
        assert torch.equal(got.values, ref.values) and torch.equal(got.row_ids, ref.row_ids) and torch.equal(got_empty, ref_empty)
        if weights:
            assert torch.allclose(got.weights, ref.weights)
    with _native.cuda_emulation():
        e, ind = sparse_fill_empty_rows(dr.SparseIds(torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64), 4), 9)
    assert e.values.tolist() == [9, 9, 9, 9] and e.row_ids.tolist() == [0, 1, 2, 3] and bool(ind.all())


def test_slice_and_reshape_emulated_kernels_match_cpu():
    g = torch.Generator().manual_seed(1)
    shape = [24, 13, 17]
    dense = (torch.rand(shape, generator=g) < 0.2) * torch.randint(1, 100, shape, generator=g)
    idx = dense.nonzero(); val = dense[dense != 0]
    for v in (val, val.float()):
        ri, rv, rs = sparse_slice(idx, v, shape, [3, 0, 5], [15, 9, 100])
        with _native.cuda_emulation():
            gi, gv, gs = sparse_slice(idx, v, shape, [3, 0, 5], [15, 9, 100])
        assert gs == rs and torch.equal(gi, ri) and torch.equal(gv, rv)
    r1, s1 = sparse_reshape(idx, shape, [13, -1, 2])
    with _native.cuda_emulation():
        g1, s2 = sparse_reshape(idx, shape, [13, -1, 2])
    assert s1 == s2 and torch.equal(g1, r1)


@pytest.mark.parametrize("mode", ["sum", "mean", "sqrtn"])
def test_sparse_segment_reductions_emulated_kernels_match_cpu(mode):
    g = torch.Generator().manual_seed(2)
    fn = {"sum": sparse_segment_sum, "mean": sparse_segment_mean, "sqrtn": sparse_segment_sqrt_n}[mode]
    data = torch.randn(300, 48, generator=g)
    indices = torch.randint(0, 300, (2000,), generator=g)
    seg = torch.sort(torch.randint(0, 70, (2000,), generator=g)).values
    d_ref = data.clone().requires_grad_(True); d_emu = data.clone().requires_grad_(True)
    ref = fn(d_ref, indices, seg, 71)
    w = torch.randn(71, 48, generator=g)
    (ref * w).sum().backward()
    with _native.cuda_emulation():
        got = fn(d_emu, indices, seg, 71)
        (got * w).sum().backward()
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4)
    assert torch.allclose(d_emu.grad, d_ref.grad, atol=1e-3, rtol=1e-3)
