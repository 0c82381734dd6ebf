"""The MonitoredTrainingSession / Estimator analogue: `Trainer` drives a zoo model with periodic full + incremental checkpoints, auto
micro-batching, a stop hook, streaming AUC -- then a second Trainer on the same directory resumes from the checkpoint chain."""
import tempfile

import torch

import _path  # noqa: F401  (repository root on sys.path)
import deeprec_b200 as dr
from deeprec_b200.data import criteo_batch
from deeprec_b200.models.zoo import build_model
from deeprec_b200.utils.trainer import Trainer

CARDS = [100] * 26


class StopAtLoss:
    def after_step(self, trainer, step, loss):
        return loss < 0.05                                      # True requests a stop (StopAtStepHook / early stopping)


def batches(n, seed):
    for s in range(n):
        yield criteo_batch(256, 13, CARDS, seed=seed + s)


def make():
    dr.embedding_variable.clear_registry()
    torch.manual_seed(0)
    model = build_model("wdl", device="cpu", cardinalities=CARDS)
    return model, dr.optim.AdagradOptimizer(model, lr=0.05)


ckpt = tempfile.mkdtemp()
loss_fn = lambda m, b: m.loss(*b)                                # noqa: E731
model, opt = make()
t = Trainer(model, opt, loss_fn, checkpoint_dir=ckpt, save_checkpoint_steps=10, save_incremental_checkpoint_steps=3, log_every_n_steps=10, micro_batch_num=2,
            hooks=[StopAtLoss()])
step = t.fit(batches(24, seed=0))
print("first run stopped at global step", step)

model2, opt2 = make()                                            # a fresh process would do exactly this
t2 = Trainer(model2, opt2, loss_fn, checkpoint_dir=ckpt, log_every_n_steps=10)
print("resumed at global step", int(opt2.global_step), "(last full checkpoint + the incremental ones after it)")
assert int(opt2.global_step) >= 20
t2.fit(batches(5, seed=100))
