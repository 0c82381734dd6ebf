"""Admission (CounterFilter / CBFFilter) and eviction (GlobalStepEvict / L2WeightEvict, applied at save time)."""
import tempfile

import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.checkpoint import Saver

counter = dr.get_embedding_variable("ad_id", 4, ev_option=dr.EmbeddingVariableOption(
    filter_option=dr.CounterFilter(filter_freq=3), evict_option=dr.GlobalStepEvict(steps_to_live=2),
    init_option=dr.InitializerOption(default_value_no_permission=0.0)))
bloom = dr.get_embedding_variable("query", 4, ev_option=dr.EmbeddingVariableOption(
    filter_option=dr.CBFFilter(filter_freq=3, max_element_size=10000, false_positive_probability=0.01, counter_type=torch.int16),
    evict_option=dr.L2WeightEvict(l2_weight_threshold=1e-9)))
opt = dr.optim.GradientDescentOptimizer([], [counter, bloom], lr=0.1)

for step in range(3):
    for ev in (counter, bloom):
        rows = ev.lookup(torch.tensor([5]))
        print(f"step {step} {ev.name}: admitted rows {ev.total_count()}, value {rows[0, :2].tolist()}")
        rows.sum().backward()
    opt.step()
assert counter.total_count() == 1 and bloom.total_count() == 1        # third occurrence admitted both

for step in range(4):                                                   # id 5 goes stale, id 9 stays fresh
    counter.lookup(torch.tensor([9])).sum().backward(); opt.step()
print("before save:", counter.total_count(), "rows (eviction only runs inside save)")
with tempfile.TemporaryDirectory() as d:
    Saver(embedding_variables=[counter, bloom], optimizer=opt).save(d + "/model.ckpt")
print("after save :", counter.total_count(), "rows; id 5 version", counter.get_version(torch.tensor([5])).item())
assert counter.get_version(torch.tensor([5])).item() == -1
