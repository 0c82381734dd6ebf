"""TensorBoard summaries as a Trainer hook (the reference: tf.summary scalars in the modelzoo -- loss, global_step/sec -- plus the
EmbeddingVariable introspection ops and ``WorkQueue.add_summary()``): loss and steps/s every ``every_n_steps``, and per EmbeddingVariable
the number of admitted rows, tracked keys, and -- for tiered tables -- the cache hit rate."""
from __future__ import annotations

import time
from typing import Optional


class SummaryHook:
    def __init__(self, logdir: str, every_n_steps: int = 100, work_queue=None):
        from torch.utils.tensorboard import SummaryWriter
        self.writer = SummaryWriter(logdir)
        self.every, self.wq = max(1, every_n_steps), work_queue
        self._t: Optional[float] = None
        self._last_step = 0

    def begin(self, trainer) -> None:
        self._t, self._last_step = time.time(), int(trainer.opt.global_step)

    def after_step(self, trainer, step: int, loss: float) -> bool:
        if step % self.every:
            return False
        w, now = self.writer, time.time()
        w.add_scalar("loss", loss, step)
        if self._t is not None and step > self._last_step:
            w.add_scalar("global_step/sec", (step - self._last_step) / max(now - self._t, 1e-9), step)
        self._t, self._last_step = now, step
        for ev in getattr(trainer.opt, "evs", []):
            w.add_scalar(f"embedding_variable/{ev.name}/rows", ev.total_count(), step)
            t = ev.table
            if hasattr(t, "total_keys"):
                w.add_scalar(f"embedding_variable/{ev.name}/tracked_keys", t.total_keys(), step)
            if hasattr(t, "cache_stats"):
                w.add_scalar(f"embedding_variable/{ev.name}/cache_hit_rate", float(t.cache_stats().get("hit_rate", 0.0)), step)
        if self.wq is not None:
            for k, v in self.wq.add_summary().items():
                w.add_scalar(f"work_queue/{k}", v, step)
        return False

    def end(self, trainer, step: int) -> None:
        self.writer.flush()
        self.writer.close()
