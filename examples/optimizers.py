"""DeepRec's optimizers on an EmbeddingVariable + a dense layer: AdagradDecay (accumulator decayed every N global steps), AdamAsync
(per-apply beta powers, optional sparse RMSProp form) and AdamW -- the sparse rows and their slot variables share ONE row of the table."""
import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr


def run(name, make):
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    ev = dr.get_embedding_variable(f"item_{name}", embedding_dim=8)
    head = torch.nn.Linear(8, 1)
    opt = make(list(head.parameters()), [ev])
    ids, y = torch.randint(0, 50, (256,)), torch.rand(256)
    first = None
    for step in range(30):
        loss = torch.nn.functional.mse_loss(head(ev.lookup(ids)).squeeze(-1), y)
        first = first if first is not None else float(loss.detach())
        opt.zero_grad(); loss.backward(); opt.step()
    print(f"{name:14s} loss {first:.4f} -> {float(loss.detach()):.4f}   rows {ev.total_count()}   slots per row: {opt.slot_names}")
    assert float(loss.detach()) < first


run("AdagradDecay", lambda p, evs: dr.optim.AdagradDecayOptimizer(p, evs, lr=0.1, accumulator_decay_step=10, accumulator_decay_rate=0.9))
run("AdamAsync", lambda p, evs: dr.optim.AdamAsyncOptimizer(p, evs, lr=0.02))
run("AdamAsync/rms", lambda p, evs: dr.optim.AdamAsyncOptimizer(p, evs, lr=0.02, apply_sparse_rmsprop=True))
run("AdamW", lambda p, evs: dr.optim.AdamWOptimizer(p, evs, lr=0.02, weight_decay=0.01))
