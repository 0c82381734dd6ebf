"""Auto graph fusion over the nn.Module tree: templates, parameter identity, numerics, idempotence."""
import torch
import torch.nn as nn

from deeprec_b200 import graph_optimizer as go
from deeprec_b200.nn import FusedMLP


def _net():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(13, 64), nn.ReLU(), nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 8), nn.LayerNorm(8), nn.GELU(approximate="tanh"),
                         nn.Linear(8, 16), nn.ReLU(), nn.Sigmoid(), nn.Linear(16, 1))


def test_linear_relu_chains_layernorm_gelu_are_fused_with_same_parameters():
    m = _net()
    params_before = {id(p) for p in m.parameters()}
    x = torch.randn(32, 13)
    ref = m(x)
    rep = go.optimize(m)
    assert rep.count("LinearReluChain") == 2 and rep.count("LayerNorm") == 1 and rep.count("Gelu") == 1
    kinds = [type(c).__name__ for c in m.children()]
    assert kinds == ["FusedMLP", "FusedLayerNorm", "FusedGelu", "FusedMLP", "Sigmoid", "Linear"]        # the 1-wide logit layer stays plain
    assert m[0].sizes == [64, 32, 8] and m[0].last_act is False and m[3].last_act is True
    assert {id(p) for p in m.parameters()} == params_before             # optimizers / checkpoints keep working
    assert torch.allclose(m(x), ref, atol=1e-5)
    # gradients flow to the original parameters
    m(x).sum().backward()
    assert all(p.grad is not None for p in m.parameters())
    assert go.optimize(m).count() == 0                                   # idempotent


def test_batchnorm_folding_only_in_eval_and_exposes_longer_chains():
    torch.manual_seed(1)
    m = nn.Sequential(nn.Linear(10, 16), nn.BatchNorm1d(16), nn.ReLU(), nn.Linear(16, 8), nn.BatchNorm1d(8))
    x = torch.randn(64, 10)
    for _ in range(3):
        m(x)                                                            # populate running statistics
    assert go.optimize(m, go.OptimizerOptions(fuse_mlp=False)).count("LinearBatchNormFold") == 0     # training mode: untouched
    m.eval()
    ref = m(x)
    rep = go.optimize(m)
    assert rep.count("LinearBatchNormFold") == 2 and rep.count("LinearReluChain") == 1
    assert len(m) == 1 and isinstance(m[0], FusedMLP) and m[0].sizes == [16, 8]
    assert torch.allclose(m(x), ref, atol=1e-5)


def test_nested_modules_dice_and_options():
    class Tower(nn.Module):
        def __init__(self):
            super().__init__()
            self.body = nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 6))
            self.act = go.Dice(6)
            self.norm = nn.LayerNorm(6)
            self.heads = nn.ModuleList([nn.Sequential(nn.Linear(6, 8), nn.ReLU(), nn.Linear(8, 1)) for _ in range(2)])

        def forward(self, x):
            h = self.norm(self.act(self.body(x)))
            return torch.cat([hd(h) for hd in self.heads], -1)

    torch.manual_seed(2)
    t = Tower()
    x = torch.randn(16, 6)
    t.train(); t(x); t.eval()
    with torch.no_grad():
        t.act.alpha.fill_(0.25)
    ref = t(x)
    assert go.optimize(t, go.OptimizerOptions(do_op_fusion=False)).count() == 0
    rep = go.optimize(t)
    assert rep.count("Dice") == 1 and rep.count("LayerNorm") == 1 and rep.count("LinearReluChain") == 3, str(rep)
    assert isinstance(t.act, go.FusedDice) and isinstance(t.heads[1][0], FusedMLP)
    assert torch.allclose(t(x), ref, atol=1e-5)


def test_zoo_model_is_rewritten_and_matches():
    from deeprec_b200.models.zoo import build_model
    torch.manual_seed(3)
    model = build_model("masknet", device="cpu")
    g = torch.Generator().manual_seed(1)
    dense = torch.randn(8, 13, generator=g); ids = torch.randint(0, 50, (26, 8), generator=g)
    model.eval()
    with torch.no_grad():
        ref = model(dense, ids)
        rep = go.optimize(model)
        assert rep.count() > 0
        assert torch.allclose(model(dense, ids), ref, atol=1e-5)
