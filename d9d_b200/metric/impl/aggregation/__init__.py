"""Sum and weighted-mean metrics (reference ``d9d/metric/impl/aggregation``)."""

from __future__ import annotations

from typing import Any

import torch

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.metric.abc import Metric
from d9d_b200.metric.component import MetricAccumulator, sync_accumulators


class SumMetric(Metric[torch.Tensor]):
    def __init__(self) -> None:
        self._acc = MetricAccumulator(torch.zeros((), dtype=torch.float32))

    def update(self, value: torch.Tensor) -> None:
        self._acc.update(value.sum())

    def sync(self, dist_context: DistributedContext) -> None:
        self._acc.sync()

    def compute(self) -> torch.Tensor:
        return self._acc.value

    def reset(self) -> None:
        self._acc.reset()

    def to(self, device: str | torch.device | int) -> None:
        self._acc.to(device)

    def state_dict(self) -> dict[str, Any]:
        return {"accumulator": self._acc.state_dict()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._acc.load_state_dict(state_dict["accumulator"])


class WeightedMeanMetric(Metric[torch.Tensor]):
    """``sum(values * weights) / sum(weights)``; both sums travel in one all-reduce."""

    def __init__(self) -> None:
        self._value = MetricAccumulator(torch.zeros((), dtype=torch.float32))
        self._weight = MetricAccumulator(torch.zeros((), dtype=torch.float32))

    def update(self, values: torch.Tensor, weights: torch.Tensor) -> None:
        self._value.update((values * weights).sum())
        self._weight.update(weights.sum())

    def sync(self, dist_context: DistributedContext) -> None:
        sync_accumulators([self._value, self._weight])

    def compute(self) -> torch.Tensor:
        return self._value.value / self._weight.value

    @property
    def accumulated_weight(self) -> torch.Tensor:
        return self._weight.value

    def reset(self) -> None:
        self._value.reset()
        self._weight.reset()

    def to(self, device: str | torch.device | int) -> None:
        self._value.to(device)
        self._weight.to(device)

    def state_dict(self) -> dict[str, Any]:
        return {"value": self._value.state_dict(), "weight": self._weight.state_dict()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._value.load_state_dict(state_dict["value"])
        self._weight.load_state_dict(state_dict["weight"])


__all__ = ["SumMetric", "WeightedMeanMetric"]
