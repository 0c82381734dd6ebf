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


def test_ckpt_format_transform_renames_in_the_index_only(tmp_path):
    """tools/ckpt_format_transform (reference: tensorflow/tools/embedding_variable/ckpt_format_transform.cc): a checkpoint saved under other tensor
    names restores into this model after an index-only rename; the data file is shared, not copied."""
    import json
    import os
    import pytest
    from deeprec_b200.tools import ckpt_format_transform as cft
    dr.embedding_variable.clear_registry()
    old = dr.get_embedding_variable("legacy_emb-1of1", 8, seed=5)
    opt = dr.optim.AdagradOptimizer([], [old], lr=0.1, global_step=GlobalStep())
    ids = torch.tensor([3, 9, 9, 12345])
    old.lookup(ids).sum().backward(); opt.step()
    want = old.lookup(ids).detach().clone()
    prefix = Saver(embedding_variables=[old], optimizer=opt).save(str(tmp_path / "src" / "m.ckpt"))
    names = list(BundleReader(prefix).entries)
    cfg = {"checkpoint_path_prefix": prefix, "output_file": str(tmp_path / "dst" / "m.ckpt-1.index"),
           "tensor_rename_map": {"legacy_emb-1of1-keys": "user/emb-keys"}, "prefix_rename_map": {"legacy_emb-1of1": "user/emb"}}
    (tmp_path / "cfg.json").write_text(json.dumps(cfg))
    assert cft.main([str(tmp_path / "cfg.json")]) == 0
    new_prefix = cfg["output_file"][: -len(".index")]
    r = BundleReader(new_prefix)
    assert sorted(r.entries) == sorted(n.replace("legacy_emb-1of1", "user/emb") for n in names)
    assert os.path.samefile(new_prefix + ".data", prefix + ".data")                      # bytes shared, CRCs still verify on read
    assert torch.equal(r.read("user/emb-values"), BundleReader(prefix).read("legacy_emb-1of1-values"))
    dr.embedding_variable.clear_registry()
    new = dr.get_embedding_variable("user/emb", 8, seed=99)
    opt2 = dr.optim.AdagradOptimizer([], [new], lr=0.1, global_step=GlobalStep())
    Saver(embedding_variables=[new], optimizer=opt2).restore(new_prefix)
    assert torch.allclose(new.lookup(ids).detach(), want)
    with pytest.raises(KeyError):
        cft.transform(prefix, str(tmp_path / "x.index"), {"no/such": "a"}, log=None)
    with pytest.raises(ValueError):
        cft.transform(prefix, str(tmp_path / "y.index"), {names[0]: names[1]}, log=None)          # collision with an existing tensor
