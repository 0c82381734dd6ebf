"""Trainer: the MonitoredTrainingSession analogue (python/training/monitored_session.py): step loop with hooks for
logging (LoggingTensorHook every N steps), periodic full checkpoints (save_checkpoint_secs / steps), incremental
checkpoints (save_incremental_checkpoint_secs), timeline profiling, streaming ACC/AUC."""
from __future__ import annotations

import os
import time
from typing import Callable, Iterable, Optional

import torch

from ..checkpoint.saver import IncrementalSaver, latest_checkpoint
from .health import FaultInjector, StepWatchdog
from .metrics import StreamingAccuracy, StreamingAUC
from .tracing import Timeline


class Trainer:
    def __init__(self, model: torch.nn.Module, optimizer, loss_fn: Callable, checkpoint_dir: Optional[str] = None,
                 save_checkpoint_steps: int = 0, save_checkpoint_secs: float = 0, save_incremental_checkpoint_secs: float = 0,
                 save_incremental_checkpoint_steps: int = 0, log_every_n_steps: int = 100, timeline_steps: int = 0,
                 micro_batch_num: int = 1, strategy=None, log: Callable[[str], None] = print, watchdog_timeout_s: float = 0,
                 hooks: Optional[list] = None):
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.dir = checkpoint_dir
        self.save_steps, self.save_secs = save_checkpoint_steps, save_checkpoint_secs
        self.incr_secs, self.incr_steps = save_incremental_checkpoint_secs, save_incremental_checkpoint_steps
        self.log_every, self.timeline_steps = log_every_n_steps, timeline_steps
        self.micro = max(1, micro_batch_num)
        # SessionRunHook analogue: objects with any of begin(trainer) / after_step(trainer, step, loss) / end(trainer, step);
        # after_step returning True requests a stop (StopAtStepHook, early stopping, NaN guard ...)
        self.hooks = list(hooks or [])
        self.strategy, self.log = strategy, log
        self.saver = IncrementalSaver(model, optimizer=optimizer) if checkpoint_dir else None
        self.timeline = Timeline() if timeline_steps else None
        self.auc, self.acc = StreamingAUC(), StreamingAccuracy()
        self._last_save = self._last_incr = time.time()
        self.watchdog = StepWatchdog(watchdog_timeout_s) if watchdog_timeout_s > 0 else None       # stalled step -> stacks + exit 86
        self.faults = FaultInjector(rank=strategy.rank if strategy is not None else 0)            # DEEPREC_FAULT=step=..,kind=..
        if self.saver and checkpoint_dir and latest_checkpoint(checkpoint_dir):
            step = self.saver.recover_incr_checkpoints(checkpoint_dir)        # failover: last full + replay deltas
            self.log(f"restored from {checkpoint_dir} at global step {step}")

    def train_step(self, batch) -> float:
        self.opt.zero_grad()
        if self.micro == 1:
            loss = self.loss_fn(self.model, batch)
            loss.backward()
        else:   # auto micro-batch (graph_execution_state.cc:635-729): slice the batch, accumulate gradients
            loss = 0.0
            for mb in _split(batch, self.micro):
                l = self.loss_fn(self.model, mb) / self.micro
                l.backward()
                loss = loss + l.detach()
        if self.strategy is not None and self.strategy.world_size > 1:
            self.strategy.allreduce_gradients([p for p in self.model.parameters() if p.numel()])
        self.opt.step()
        return float(loss.detach()) if torch.is_tensor(loss) else float(loss)

    def fit(self, batches: Iterable, max_steps: Optional[int] = None) -> int:
        step = int(self.opt.global_step)
        t0, n0 = time.time(), step
        for h in self.hooks:
            if hasattr(h, "begin"):
                h.begin(self)
        stop = False
        for batch in batches:
            if stop:
                break
            if max_steps is not None and step - n0 >= max_steps:
                break
            if self.timeline is not None and step - n0 < self.timeline_steps:
                with self.timeline.span(f"step_{step}", device=torch.cuda.is_available()):
                    loss = self.train_step(batch)
            else:
                loss = self.train_step(batch)
            step = int(self.opt.global_step)
            if self.watchdog is not None:
                self.watchdog.tick()
            self.faults.maybe_fail(step)
            for h in self.hooks:
                if hasattr(h, "after_step") and h.after_step(self, step, loss):
                    stop = True
            if self.log_every and step % self.log_every == 0:
                dt = time.time() - t0
                self.log(f"global_step {step}  loss {loss:.5f}  {(step - n0) / max(dt, 1e-9):.2f} global_step/sec")
            now = time.time()
            if self.saver:
                if (self.save_steps and step % self.save_steps == 0) or (self.save_secs and now - self._last_save >= self.save_secs):
                    self.saver.save(os.path.join(self.dir, "model.ckpt"), step); self._last_save = now
                elif (self.incr_steps and step % self.incr_steps == 0) or (self.incr_secs and now - self._last_incr >= self.incr_secs):
                    if latest_checkpoint(self.dir):
                        self.saver.incremental_save(os.path.join(self.dir, "model.ckpt"), step); self._last_incr = now
        if self.timeline is not None and self.dir:
            self.timeline.save(os.path.join(self.dir, "timeline.json"))
        for h in self.hooks:
            if hasattr(h, "end"):
                h.end(self, step)
        return step

    @torch.no_grad()
    def evaluate(self, batches: Iterable, predict_fn: Callable, max_steps: int = 100):
        self.model.eval()
        for i, batch in enumerate(batches):
            if i >= max_steps:
                break
            prob, label = predict_fn(self.model, batch)
            self.auc.update(prob, label); self.acc.update(prob, label)
        self.model.train()
        return {"acc": self.acc.result(), "auc": self.auc.result()}


def _split(batch, n):
    if isinstance(batch, dict):
        keys = list(batch)
        parts = {k: torch.chunk(batch[k], n, dim=0) for k in keys}
        return [{k: parts[k][i] for k in keys} for i in range(len(parts[keys[0]]))]
    dense, ids, y = batch
    return list(zip(torch.chunk(dense, n, 0), torch.chunk(ids, n, 1), torch.chunk(y, n, 0)))
