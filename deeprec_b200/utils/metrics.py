"""Streaming metrics (tf.metrics.accuracy / tf.metrics.auc(num_thresholds=1000), modelzoo/dlrm/train.py:279-286)."""
from __future__ import annotations

import torch


class StreamingAUC:
    def __init__(self, num_thresholds: int = 1000):
        self.n = num_thresholds
        self.pos = torch.zeros(num_thresholds + 1, dtype=torch.float64)
        self.neg = torch.zeros(num_thresholds + 1, dtype=torch.float64)

    def update(self, prob: torch.Tensor, label: torch.Tensor) -> None:
        p = prob.detach().float().cpu().clamp(0, 1)
        y = label.detach().float().cpu()
        b = (p * self.n).long().clamp(0, self.n)
        self.pos += torch.bincount(b[y > 0.5], minlength=self.n + 1).double()
        self.neg += torch.bincount(b[y <= 0.5], minlength=self.n + 1).double()

    def result(self) -> float:
        tp = torch.flip(torch.cumsum(torch.flip(self.pos, [0]), 0), [0])
        fp = torch.flip(torch.cumsum(torch.flip(self.neg, [0]), 0), [0])
        P, N = self.pos.sum(), self.neg.sum()
        if P == 0 or N == 0:
            return 0.5
        tpr = torch.cat([tp / P, torch.zeros(1, dtype=torch.float64)])
        fpr = torch.cat([fp / N, torch.zeros(1, dtype=torch.float64)])
        return float(torch.trapz(torch.flip(tpr, [0]), torch.flip(fpr, [0])))


class StreamingAccuracy:
    def __init__(self):
        self.correct, self.total = 0, 0

    def update(self, prob: torch.Tensor, label: torch.Tensor) -> None:
        self.correct += int(((prob.detach().float().cpu() > 0.5).float() == label.detach().float().cpu()).sum())
        self.total += label.numel()

    def result(self) -> float:
        return self.correct / max(1, self.total)
