"""DIN attention: the first-layer algebra used by the fused kernel, dispatch rules, CPU path == composite reference."""
import torch
import torch.nn as nn

from deeprec_b200.ops import attention as A


def _unit(D=32, H1=80, H2=40):
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(4 * D, H1), nn.Sigmoid(), nn.Linear(H1, H2), nn.Sigmoid(), nn.Linear(H2, 1))


def _emulate_kernel(q, k, mask, att):
    """Step-by-step emulation of csrc/cuda/attention_kernels.cu in torch (same split weights, same masking / softmax rules)."""
    D = q.shape[-1]
    Wq, Wk, Wp = A.split_first_layer(att[0].weight, D)
    kk = k * mask.unsqueeze(-1)
    hq = att[0].bias + q @ Wq.T                                            # once per sample
    h1 = torch.sigmoid(hq.unsqueeze(1) + kk @ Wk.T + (kk * q.unsqueeze(1)) @ Wp.T)
    h2 = torch.sigmoid(h1 @ att[2].weight.T + att[2].bias)
    s = h2 @ att[4].weight.view(-1) + att[4].bias
    mx = torch.where(mask, s, torch.full_like(s, float("-inf"))).max(-1, keepdim=True).values
    e = torch.where(mask, torch.exp(s - torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))), torch.zeros_like(s))
    den = e.sum(-1, keepdim=True)
    w = torch.where(den > 0, e / den.clamp_min(1e-30), torch.zeros_like(e))
    return (w.unsqueeze(-1) * kk).sum(1)


def test_kernel_algebra_matches_reference():
    att = _unit()
    g = torch.Generator().manual_seed(1)
    q = torch.randn(16, 32, generator=g); k = torch.randn(16, 50, 32, generator=g)
    lens = torch.randint(0, 51, (16,), generator=g); lens[0] = 0; lens[1] = 50       # an empty and a full history
    mask = torch.arange(50).unsqueeze(0) < lens.unsqueeze(1)
    with torch.no_grad():
        ref = A.din_attention_reference(q, k * mask.unsqueeze(-1), mask, att)
        emu = _emulate_kernel(q, k, mask, att)
    assert torch.allclose(emu, ref, atol=1e-5), (emu - ref).abs().max()
    assert torch.all(ref[0] == 0)                                           # no valid position -> zero vector


def test_dispatch_falls_back_on_cpu_and_under_autograd():
    att = _unit(D=8, H1=16, H2=8)
    q = torch.randn(4, 8, requires_grad=True); k = torch.randn(4, 6, 8); mask = torch.ones(4, 6, dtype=torch.bool)
    out = A.din_attention(q, k, mask, att)
    out.sum().backward()
    assert q.grad is not None and att[0].weight.grad is not None
    assert A._fusable(att) and not A._fusable(nn.Sequential(nn.Linear(8, 1)))


def test_composite_training_path_matches_reference_values_and_gradients():
    att_a, att_b = _unit(), _unit()
    g = torch.Generator().manual_seed(2)
    q = torch.randn(12, 32, generator=g); k = torch.randn(12, 9, 32, generator=g)
    lens = torch.randint(0, 10, (12,), generator=g); lens[0] = 0
    mask = torch.arange(9).unsqueeze(0) < lens.unsqueeze(1)
    k = k * mask.unsqueeze(-1)
    qa, ka = q.clone().requires_grad_(), k.clone().requires_grad_()
    qb, kb = q.clone().requires_grad_(), k.clone().requires_grad_()
    oa = A.din_attention_composite(qa, ka, mask, att_a)
    ob = A.din_attention_reference(qb, kb, mask, att_b)
    assert torch.allclose(oa, ob, atol=1e-5)
    w = torch.randn(12, 32, generator=g)
    (oa * w).sum().backward(); (ob * w).sum().backward()
    assert torch.allclose(qa.grad, qb.grad, atol=1e-5) and torch.allclose(ka.grad, kb.grad, atol=1e-5)
    for pa, pb in zip(att_a.parameters(), att_b.parameters()):
        assert torch.allclose(pa.grad, pb.grad, atol=1e-5)
