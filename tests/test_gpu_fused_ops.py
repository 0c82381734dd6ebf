"""Fused LayerNorm / L2-normalize / GELU / Dice kernels vs PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fused_ops_forward_backward():
    from deeprec_b200.ops.fused_ops import dice, fused_l2_normalize, fused_layer_norm, gelu
    torch.manual_seed(0)
    x = torch.randn(333, 100, device="cuda", requires_grad=True)
    g = torch.randn(100, device="cuda", requires_grad=True); b = torch.randn(100, device="cuda", requires_grad=True)
    y = fused_layer_norm(x, g, b, 1e-5); ref = torch.nn.functional.layer_norm(x, (100,), g, b, 1e-5)
    assert torch.allclose(y, ref, atol=1e-4)
    go = torch.randn_like(y)
    gx, gg, gb = torch.autograd.grad(y, (x, g, b), go); rx, rg, rb = torch.autograd.grad(ref, (x, g, b), go)
    assert torch.allclose(gx, rx, atol=1e-3) and torch.allclose(gg, rg, atol=1e-2) and torch.allclose(gb, rb, atol=1e-3)
    y = fused_l2_normalize(x); ref = torch.nn.functional.normalize(x, dim=-1)
    assert torch.allclose(y, ref, atol=1e-5)
    assert torch.allclose(torch.autograd.grad(y, x, go)[0], torch.autograd.grad(ref, x, go)[0], atol=1e-4)
    for approx in (True, False):
        y = gelu(x, approx); ref = torch.nn.functional.gelu(x, approximate="tanh" if approx else "none")
        assert torch.allclose(y, ref, atol=1e-4)
        assert torch.allclose(torch.autograd.grad(y, x, go)[0], torch.autograd.grad(ref, x, go)[0], atol=1e-3)
    xd = x.detach(); alpha = torch.rand(100, device="cuda"); m = xd.mean(0); v = xd.var(0, unbiased=False)
    p = torch.sigmoid((xd - m) * torch.rsqrt(v + 1e-9))
    assert torch.allclose(dice(xd, alpha, m, v), p * xd + (1 - p) * alpha * xd, atol=1e-4)
