from __future__ import annotations

from typing import Any

import torch

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.metric.abc import Metric
from d9d_b200.metric.component import MetricAccumulator


def histogram_auroc(pos_hist: torch.Tensor, neg_hist: torch.Tensor) -> torch.Tensor:
    """``P(score_pos > score_neg) + 0.5 P(equal bin)`` from per-bin counts; 0.5 when a class is absent.
    Branch-free (``torch.where``) so it never synchronises with the host."""
    total_pos, total_neg = pos_hist.sum(), neg_hist.sum()
    pos_above = total_pos - pos_hist.cumsum(0)
    area = (neg_hist * (pos_above + 0.5 * pos_hist)).sum()
    valid = (total_pos > 0) & (total_neg > 0)
    return torch.where(valid, area / (total_pos * total_neg).clamp_min(1e-30), pos_hist.new_tensor(0.5))


class BinaryAUROCMetric(Metric[torch.Tensor]):
    """Streaming AUROC approximation with two fixed-size score histograms (constant memory, all-reduce friendly).

    Parity: reference ``d9d/metric/impl/classification/auroc.py:10-130``; both histograms share one ``[2, bins]``
    accumulator here.
    """

    def __init__(self, num_bins: int = 10000):
        self._bins = num_bins
        self._hist = MetricAccumulator(torch.zeros(2, num_bins, dtype=torch.float32))

    def update(self, probs: torch.Tensor, labels: torch.Tensor) -> None:
        probs, labels = probs.reshape(-1), labels.reshape(-1)
        if probs.numel() != labels.numel():
            raise ValueError("Predictions and labels should have the same number of elements")
        device = self._hist.value.device
        idx = (probs.to(device) * self._bins).long().clamp_(0, self._bins - 1)
        lab = labels.to(device).float()
        batch = torch.zeros(2, self._bins, dtype=torch.float32, device=device)
        batch[0].index_add_(0, idx, lab)
        batch[1].index_add_(0, idx, 1.0 - lab)
        self._hist.update(batch)

    def sync(self, dist_context: DistributedContext) -> None:
        self._hist.sync()

    def compute(self) -> torch.Tensor:
        h = self._hist.value
        return histogram_auroc(h[0], h[1])

    def reset(self) -> None:
        self._hist.reset()

    def to(self, device: str | torch.device | int) -> None:
        self._hist.to(device)

    def state_dict(self) -> dict[str, Any]:
        return {"hist": self._hist.state_dict()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._hist.load_state_dict(state_dict["hist"])
