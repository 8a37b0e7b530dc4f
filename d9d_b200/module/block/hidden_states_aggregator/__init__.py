"""Per-layer hidden-state snapshots (e.g. masked means) carried through the model / pipeline stages.

Parity: reference ``d9d/module/block/hidden_states_aggregator``.
"""

from __future__ import annotations

import abc
import enum

import torch


class BaseHiddenStatesAggregator(abc.ABC):
    @abc.abstractmethod
    def add_hidden_states(self, hidden_states: torch.Tensor) -> None: ...

    @abc.abstractmethod
    def pack_with_snapshot(self, snapshot: torch.Tensor | None) -> torch.Tensor | None: ...


class HiddenStatesAggregatorNoOp(BaseHiddenStatesAggregator):
    def add_hidden_states(self, hidden_states: torch.Tensor) -> None:
        return None

    def pack_with_snapshot(self, snapshot: torch.Tensor | None) -> torch.Tensor | None:
        return None


class HiddenStatesAggregatorMean(BaseHiddenStatesAggregator):
    """Masked mean over the sequence dim, computed in fp32 as states arrive."""

    def __init__(self, agg_mask: torch.Tensor) -> None:
        self._mask = agg_mask
        self._collected: list[torch.Tensor] = []

    def add_hidden_states(self, hidden_states: torch.Tensor) -> None:
        mask = self._mask
        summed = (hidden_states.float() * mask.unsqueeze(-1)).sum(dim=1)
        self._collected.append((summed / mask.sum(dim=1, keepdim=True)).to(hidden_states.dtype))

    def pack_with_snapshot(self, snapshot: torch.Tensor | None) -> torch.Tensor | None:
        if not self._collected:
            return None
        stacked = torch.stack(self._collected, dim=0)
        self._collected.clear()
        return stacked if snapshot is None else torch.cat([snapshot, stacked], dim=0)


class HiddenStatesAggregationMode(enum.StrEnum):
    no = "no"
    mean = "mean"


def create_hidden_states_aggregator(mode: HiddenStatesAggregationMode, agg_mask: torch.Tensor | None) -> BaseHiddenStatesAggregator:
    if mode == HiddenStatesAggregationMode.no:
        return HiddenStatesAggregatorNoOp()
    if mode == HiddenStatesAggregationMode.mean:
        if agg_mask is None:
            raise ValueError("You have to specify aggregation mask")
        return HiddenStatesAggregatorMean(agg_mask)
    raise ValueError("Unknown hidden states aggregation mode")


__all__ = ["BaseHiddenStatesAggregator", "HiddenStatesAggregationMode", "create_hidden_states_aggregator"]
