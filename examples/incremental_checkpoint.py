"""Full + incremental checkpoints, chain recovery, and restoring 1 shard into 2 partitions."""
import tempfile

import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.checkpoint import IncrementalSaver, Saver
from deeprec_b200.optim import GlobalStep


def make(tag):
    dr.embedding_variable.clear_registry()
    ev = dr.get_embedding_variable("emb", 8, seed=1)
    head = torch.nn.Linear(8, 1)
    return ev, head, dr.optim.AdamOptimizer(head.parameters(), [ev], lr=0.01, global_step=GlobalStep())


def train(ev, head, opt, lo, hi, steps, seed):
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        ids = torch.randint(lo, hi, (64,), generator=g)
        loss = (head(ev.lookup(ids)).squeeze(-1) - 1).pow(2).mean()
        opt.zero_grad(); loss.backward(); opt.step()


with tempfile.TemporaryDirectory() as d:
    ev, head, opt = make("a")
    sv = IncrementalSaver(torch.nn.ModuleList([ev, head]), optimizer=opt)
    train(ev, head, opt, 0, 500, 5, 0); print("full       :", sv.save(d + "/model.ckpt"))
    train(ev, head, opt, 400, 600, 3, 1); print("incremental:", sv.incremental_save(d + "/model.ckpt"))
    train(ev, head, opt, 550, 700, 3, 2); print("incremental:", sv.incremental_save(d + "/model.ckpt"))
    probe = torch.arange(0, 700)
    want = ev.table.lookup(probe).clone()

    ev2, head2, opt2 = make("b")                        # a fresh process after a crash
    step = IncrementalSaver(torch.nn.ModuleList([ev2, head2]), optimizer=opt2).recover_incr_checkpoints(d)
    print("recovered at step", step, "rows", ev2.total_count())
    assert step == 11 and torch.equal(ev2.table.lookup(probe), want)

    prefix = Saver(torch.nn.ModuleList([ev, head]), optimizer=opt).save(d + "/full.ckpt")
    total = 0
    for part in range(2):                               # 1 -> 2 re-sharding on restore (key % 1000 % 2)
        evp, headp, optp = make("p")
        Saver(torch.nn.ModuleList([evp, headp]), optimizer=optp, partition_id=part, partition_num=2).restore(prefix)
        print(f"partition {part}: {evp.total_count()} rows")
        total += evp.total_count()
    assert total == ev.total_count()
