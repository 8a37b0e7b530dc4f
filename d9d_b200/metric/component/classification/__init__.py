"""Confusion-matrix building blocks: processors -> accumulator -> statistic -> aggregation."""

from __future__ import annotations

import dataclasses
import enum
from typing import Any, Protocol

import torch
import torch.nn.functional as F
from torch.distributed.checkpoint.stateful import Stateful

from d9d_b200.metric.component.accumulator import MetricAccumulator


@dataclasses.dataclass(kw_only=True, slots=True)
class ConfusionMatrix:
    tp: torch.Tensor
    fp: torch.Tensor
    tn: torch.Tensor
    fn: torch.Tensor


# ----------------------------------------------------------------------------- processors
class ClassificationPredictionsProcessor(Protocol):
    def __call__(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]: ...


class TopKProcessor(ClassificationPredictionsProcessor):
    """Hit/miss of the target inside the top-k logits -> ``(hit[..., 1], ones[..., 1])``."""

    def __init__(self, k: int) -> None:
        self._k = k

    def __call__(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        top = torch.topk(preds, self._k, dim=-1).indices
        hit = (top == targets.unsqueeze(-1)).any(dim=-1, keepdim=True).long()
        return hit, torch.ones_like(hit)


class OneHotProcessor(ClassificationPredictionsProcessor):
    """argmax predictions and integer / one-hot targets -> one-hot ``[..., C]`` pairs."""

    def __init__(self, num_classes: int) -> None:
        self._c = num_classes

    def __call__(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if preds.shape[-1] != self._c:
            raise ValueError(f"Expected last dimension of preds to equal num_classes={self._c}, got {preds.shape[-1]}")
        p = F.one_hot(preds.argmax(dim=-1), num_classes=self._c).long()
        if targets.shape == preds.shape:
            t = targets.long()
        elif targets.shape == preds.shape[:-1]:
            t = F.one_hot(targets.long(), num_classes=self._c).long()
        elif targets.shape == (*preds.shape[:-1], 1):
            t = F.one_hot(targets.squeeze(-1).long(), num_classes=self._c).long()
        else:
            raise ValueError(f"Targets shape {targets.shape} is incompatible with predictions shape {preds.shape}. "
                             "Expected shape to be (...), (..., 1), or (..., C).")
        return p, t


class ThresholdProcessor(ClassificationPredictionsProcessor):
    """Binarise probabilities at ``threshold`` (1-D inputs get a trailing dim)."""

    def __init__(self, threshold: float) -> None:
        self._threshold = threshold

    def __call__(self, preds: torch.Tensor, targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if preds.ndim == 1:
            preds = preds.unsqueeze(-1)
        if targets.ndim == 1:
            targets = targets.unsqueeze(-1)
        return (preds > self._threshold).float(), targets.float()


# ----------------------------------------------------------------------------- accumulator
class ConfusionMatrixAccumulator(Stateful):
    """tp/fp/tn/fn per output kept in ONE ``[4, C]`` accumulator (one all-reduce per sync instead of four)."""

    def __init__(self, num_outputs: int):
        self._n = num_outputs
        self._counts = MetricAccumulator(torch.zeros(4, num_outputs, dtype=torch.long))

    @property
    def state(self) -> ConfusionMatrix:
        v = self._counts.value
        return ConfusionMatrix(tp=v[0], fp=v[1], tn=v[2], fn=v[3])

    def update(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        if preds.shape != targets.shape:
            raise ValueError(f"preds and targets must have same shape, got {preds.shape} and {targets.shape}")
        if preds.shape[-1] != self._n:
            raise ValueError(f"Expected {self._n} outputs, got {preds.shape[-1]}")
        p = preds.long().reshape(-1, self._n)
        t = targets.long().reshape(-1, self._n)
        tp = (p * t).sum(0)
        pos_pred, pos_true, total = p.sum(0), t.sum(0), p.shape[0]
        fp, fn = pos_pred - tp, pos_true - tp
        tn = total - tp - fp - fn
        self._counts.update(torch.stack([tp, fp, tn, fn]))

    def sync(self) -> None:
        self._counts.sync()

    def reset(self) -> None:
        self._counts.reset()

    def to(self, device: str | torch.device | int) -> None:
        self._counts.to(device)

    def state_dict(self) -> dict[str, Any]:
        return {"counts": self._counts.state_dict()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._counts.load_state_dict(state_dict["counts"])


# ----------------------------------------------------------------------------- statistics
class ConfusionMatrixStatistic(Protocol):
    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor: ...


class PrecisionStatistic(ConfusionMatrixStatistic):
    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor:
        return matrix.tp / (matrix.tp + matrix.fp)


class RecallStatistic(ConfusionMatrixStatistic):
    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor:
        return matrix.tp / (matrix.tp + matrix.fn)


class FBetaStatistic(ConfusionMatrixStatistic):
    def __init__(self, beta: float) -> None:
        self._b2 = beta**2

    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor:
        scaled_tp = (1 + self._b2) * matrix.tp
        return scaled_tp / (scaled_tp + self._b2 * matrix.fn + matrix.fp)


class AccuracyStatistic(ConfusionMatrixStatistic):
    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor:
        return (matrix.tp + matrix.tn) / (matrix.tp + matrix.tn + matrix.fp + matrix.fn)


# ----------------------------------------------------------------------------- aggregation
class ClassificationAggregationMethod(enum.StrEnum):
    MICRO = "micro"  # statistic of the summed matrix
    MACRO = "macro"  # unweighted mean of per-class statistics
    WEIGHTED = "weighted"  # support-weighted mean of per-class statistics
    NONE = "none"  # per-class vector


class ConfusionMatrixAggregator:
    def __init__(self, method: ClassificationAggregationMethod, statistic: ConfusionMatrixStatistic) -> None:
        self._method, self._statistic = method, statistic

    def __call__(self, matrix: ConfusionMatrix) -> torch.Tensor:
        if self._method == ClassificationAggregationMethod.MICRO:
            return self._statistic(ConfusionMatrix(tp=matrix.tp.sum(), fp=matrix.fp.sum(), tn=matrix.tn.sum(), fn=matrix.fn.sum()))
        scores = self._statistic(matrix)
        if self._method == ClassificationAggregationMethod.MACRO:
            return scores.mean()
        if self._method == ClassificationAggregationMethod.WEIGHTED:
            support = matrix.tp + matrix.fn
            return (scores * support).sum() / support.sum()
        if self._method == ClassificationAggregationMethod.NONE:
            return scores
        raise ValueError(f"Unknown aggregation method: {self._method}")


__all__ = [
    "AccuracyStatistic",
    "ClassificationAggregationMethod",
    "ClassificationPredictionsProcessor",
    "ConfusionMatrix",
    "ConfusionMatrixAccumulator",
    "ConfusionMatrixAggregator",
    "ConfusionMatrixStatistic",
    "FBetaStatistic",
    "OneHotProcessor",
    "PrecisionStatistic",
    "RecallStatistic",
    "ThresholdProcessor",
    "TopKProcessor",
]
