"""EmbeddingVariable basics: create, train, introspect, export."""
import argparse

import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr

p = argparse.ArgumentParser(); p.add_argument("--device", default="cpu"); a = p.parse_args()
storage = dr.StorageType.HBM if a.device.startswith("cuda") else dr.StorageType.DRAM
ev = dr.get_embedding_variable("user_id", embedding_dim=8, device=a.device,
                               ev_option=dr.EmbeddingVariableOption(storage_option=dr.StorageOption(storage),
                                                                    init_option=dr.InitializerOption(default_value_dim=1024)))
head = torch.nn.Linear(8, 1, device=a.device)
opt = dr.optim.AdagradOptimizer(head.parameters(), [ev], lr=0.1)

ids = torch.tensor([3, 3, 10**12, 42], device=a.device)          # any int64 id, no vocabulary
assert ev.total_count() == 0                                      # lookups never create rows
for step in range(5):
    loss = (head(ev.lookup(ids)).squeeze(-1) - 1.0).pow(2).mean()
    opt.zero_grad(); loss.backward(); opt.step()
    print(f"step {step} loss {loss.item():.4f}")

uniq = torch.tensor([3, 10**12, 42, 7], device=a.device)
print("rows:", ev.total_count(), "frequency:", ev.get_frequency(uniq).tolist(), "version:", ev.get_version(uniq).tolist())
keys, values, versions, freqs = ev.export()
print("exported", keys.numel(), "keys; slot 'accumulator' of id 3:", ev.slot_values(torch.tensor([3], device=a.device), "accumulator")[0, :3].tolist())
assert ev.total_count() == 3 and ev.get_frequency(uniq).tolist() == [10, 5, 5, 0]
