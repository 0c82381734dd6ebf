"""Checkpoint tools: filtered-feature shrink, low-precision conversion."""
import torch
from torch import nn

import deeprec_b200 as dr
from deeprec_b200.checkpoint import BundleReader, Saver
from deeprec_b200.optim import GlobalStep
from deeprec_b200.tools.low_precision_optimize import convert, load_tensor
from deeprec_b200.tools.shrink_ckpt import shrink


class _M(nn.Module):
    def __init__(self):
        super().__init__()
        self.ev = dr.get_embedding_variable("tools/emb", 16, ev_option=dr.EmbeddingVariableOption(filter_option=dr.CounterFilter(3)))
        self.fc = nn.Linear(16, 8)


def test_shrink_and_low_precision(tmp_path):
    m = _M()
    opt = dr.optim.AdagradOptimizer(m, lr=0.1, global_step=GlobalStep())
    for _ in range(2):
        m.fc(m.ev.lookup(torch.cat([torch.arange(50), torch.arange(10)]))).sum().backward(); opt.step()
    prefix = Saver(m, optimizer=opt).save(str(tmp_path / "m.ckpt"))
    r = BundleReader(prefix)
    assert r.read("tools/emb-keys_filtered").numel() > 0
    st = shrink(prefix, str(tmp_path / "small"))
    assert st["dropped"] == 4 and not BundleReader(str(tmp_path / "small")).has("tools/emb-keys_filtered")
    ref = r.read("tools/emb-values")
    for dt, tol, ratio in (("bf16", 1e-2, 1.0), ("int8", 2e-2, 1.0)):
        info = convert(prefix, str(tmp_path / dt), dt)
        assert info["ratio"] < ratio
        got = load_tensor(BundleReader(str(tmp_path / dt)), "tools/emb-values")
        assert (got - ref).abs().max().item() < tol * (ref.abs().max().item() + 1e-6) + 1e-3


def test_inspect_checkpoint_lists_tensors_and_summarises_evs(tmp_path, capsys):
    import torch
    import deeprec_b200 as dr
    from deeprec_b200.checkpoint import Saver
    from deeprec_b200.optim import GlobalStep
    from deeprec_b200.tools import inspect_checkpoint
    dr.embedding_variable.clear_registry()
    ev = dr.get_embedding_variable("insp/emb", 8, ev_option=dr.EmbeddingVariableOption(filter_option=dr.CounterFilter(2)), seed=2)
    opt = dr.optim.AdamOptimizer([], [ev], lr=0.01, global_step=GlobalStep())
    for ids in ([1, 2, 3, 3], [3, 4, 1001, 1001]):
        ev.lookup(torch.tensor(ids)).sum().backward(); opt.step()
    prefix = Saver(embedding_variables=[ev], optimizer=opt).save(str(tmp_path / "m.ckpt"))
    assert inspect_checkpoint.main([prefix]) == 0
    listing = capsys.readouterr().out
    assert "insp/emb-keys" in listing and "insp/emb-keys_filtered" in listing and "tensors" in listing
    assert inspect_checkpoint.main([str(tmp_path), "--ev", "--partitions", "2"]) == 0          # directory -> latest checkpoint
    summ = capsys.readouterr().out
    assert "insp/emb: keys=2 dim=8 filtered_keys=3 slots=['m', 'v']" in summ and "rows_per_partition=[0, 2]" in summ, summ
    assert inspect_checkpoint.main([prefix, "--tensor", "insp/emb-keys"]) == 0
    assert "shape=(2,)" in capsys.readouterr().out
