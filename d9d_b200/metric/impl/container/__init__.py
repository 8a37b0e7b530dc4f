from __future__ import annotations

from collections.abc import Mapping
from typing import Any

import torch

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.metric.abc import Metric


class ComposeMetric(Metric[dict[str, Any]]):
    """Named collection of metrics behaving as one (children are updated individually)."""

    def __init__(self, children: Mapping[str, Metric]):
        self._children = children

    def update(self, *args: Any, **kwargs: Any) -> None:
        raise ValueError("Cannot update ComposeMetric directly - you can only update its children")

    def __getitem__(self, item: str) -> Metric:
        return self._children[item]

    @property
    def children(self) -> Mapping[str, Metric]:
        return self._children

    def sync(self, dist_context: DistributedContext) -> None:
        for m in self._children.values():
            m.sync(dist_context)

    def compute(self) -> dict[str, Any]:
        return {k: m.compute() for k, m in self._children.items()}

    def reset(self) -> None:
        for m in self._children.values():
            m.reset()

    def to(self, device: str | torch.device | int) -> None:
        for m in self._children.values():
            m.to(device)

    def state_dict(self) -> dict[str, Any]:
        return {k: m.state_dict() for k, m in self._children.items()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        for k, m in self._children.items():
            m.load_state_dict(state_dict[k])


__all__ = ["ComposeMetric"]
