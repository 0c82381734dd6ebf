"""csrc/cuda/attention_kernels.cu (fused DIN attention forward and backward) on the CUDA-on-CPU emulation against autograd through the
composite fp32 reference -- the CPU twin of tests/test_gpu_zz_attention*.py / test_gpu_zzzz_din_attention_bwd.py."""
import os

import pytest
import torch
import torch.nn as nn

from deeprec_b200 import _native

pytestmark = [pytest.mark.timeout(900)]


QUICK = os.environ.get("DEEPREC_EMU_QUICK") == "1"          # the sanitizer re-runs (tests/test_cuda_emu_sanitizers.py) use the small case only


@pytest.mark.parametrize("B,L,D,H1,H2", [(7, 9, 16, 24, 12)] if QUICK else [(5, 50, 32, 80, 40), (9, 20, 32, 80, 40), (7, 9, 16, 24, 12)])
def test_fused_din_attention_forward_and_gradients_on_the_emulation(B, L, D, H1, H2):
    from deeprec_b200.ops.attention import din_attention, din_attention_fused_train, din_attention_reference
    torch.manual_seed(B + L)
    att = nn.Sequential(nn.Linear(4 * D, H1), nn.Sigmoid(), nn.Linear(H1, H2), nn.Sigmoid(), nn.Linear(H2, 1))
    q0, k0 = torch.randn(B, D) * 0.5, torch.randn(B, L, D) * 0.5
    mask = torch.rand(B, L) < 0.7
    mask[1] = False                                              # a sample without any valid history position
    g = torch.randn(B, D)
    res = []
    for fn in (din_attention_reference, din_attention_fused_train):
        att.zero_grad()
        q, k = q0.clone().requires_grad_(True), k0.clone().requires_grad_(True)
        if fn is din_attention_reference:
            out = fn(q, k * mask.unsqueeze(-1), mask, att)       # the fused kernels mask k themselves
            out.backward(g)
        else:
            with _native.cuda_emulation():
                out = fn(q, k, mask, att)
                out.backward(g)
        res.append([out.detach(), q.grad, k.grad] + [p.grad.clone() for p in att.parameters()])
    names = ["out", "dq", "dk", "dW1", "db1", "dW2", "db2", "dW3", "db3"]
    for n, a, b in zip(names, *res):
        scale = max(1.0, float(a.abs().max()))
        assert torch.allclose(a, b, atol=2e-3 * scale, rtol=2e-3), (n, float((a - b).abs().max()), scale)
    with torch.no_grad(), _native.cuda_emulation():                     # the inference dispatch (forward kernel only)
        inf = din_attention(q0, k0 * mask.unsqueeze(-1), mask, att)
    assert torch.allclose(inf, res[0][0], atol=2e-3, rtol=2e-3)
