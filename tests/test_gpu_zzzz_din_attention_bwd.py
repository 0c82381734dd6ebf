"""Fused DIN attention TRAINING path (csrc/cuda/attention_kernels.cu::k_din_attention_bwd): gradients of q, k and of all six parameters of
the attention unit against autograd through the composite fp32 reference.  Written after the round's GPU budget was spent: sorts late."""
import pytest
import torch
import torch.nn as nn

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.mark.parametrize("B,L,D,H1,H2", [(64, 50, 32, 80, 40), (300, 20, 32, 80, 40), (7, 9, 16, 24, 12)])
def test_fused_din_attention_gradients_match_autograd(B, L, D, H1, H2):
    from deeprec_b200.ops.attention import din_attention_fused_train, din_attention_reference
    torch.manual_seed(B + L)
    dev = "cuda"
    att = nn.Sequential(nn.Linear(4 * D, H1), nn.Sigmoid(), nn.Linear(H1, H2), nn.Sigmoid(), nn.Linear(H2, 1)).to(dev)
    q0, k0 = torch.randn(B, D, device=dev) * 0.5, torch.randn(B, L, D, device=dev) * 0.5
    mask = torch.rand(B, L, device=dev) < 0.7
    mask[1] = False                                              # a sample without any valid history position
    g = torch.randn(B, D, device=dev)
    res = []
    for fn in (din_attention_reference, din_attention_fused_train):
        att.zero_grad()
        q, k = q0.clone().requires_grad_(True), k0.clone().requires_grad_(True)
        km = k * mask.unsqueeze(-1) if fn is din_attention_reference else k     # the fused kernels mask k themselves
        out = fn(q, km, mask, att)
        out.backward(g)
        res.append([out.detach(), q.grad, k.grad] + [p.grad.clone() for p in att.parameters()])
    names = ["out", "dq", "dk", "dW1", "db1", "dW2", "db2", "dW3", "db3"]
    for n, a, b in zip(names, *res):
        scale = max(1.0, float(a.abs().max()))
        assert torch.allclose(a, b, atol=2e-3 * scale, rtol=2e-3), (n, float((a - b).abs().max()), scale)
