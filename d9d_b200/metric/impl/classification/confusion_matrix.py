from __future__ import annotations

from typing import Any, Self

import torch

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.metric.abc import Metric
from d9d_b200.metric.component.classification import (
    AccuracyStatistic,
    ClassificationAggregationMethod,
    ClassificationPredictionsProcessor,
    ConfusionMatrixAccumulator,
    ConfusionMatrixAggregator,
    ConfusionMatrixStatistic,
    FBetaStatistic,
    OneHotProcessor,
    PrecisionStatistic,
    RecallStatistic,
    ThresholdProcessor,
    TopKProcessor,
)


class ConfusionMatrixMetric(Metric[torch.Tensor]):
    """processor -> confusion-matrix accumulator -> (aggregated) statistic."""

    def __init__(self, processor: ClassificationPredictionsProcessor, accumulator: ConfusionMatrixAccumulator,
                 aggregator: ConfusionMatrixAggregator) -> None:
        self._processor, self._accumulator, self._aggregator = processor, accumulator, aggregator

    def update(self, preds: torch.Tensor, targets: torch.Tensor) -> None:
        self._accumulator.update(*self._processor(preds, targets))

    def sync(self, dist_context: DistributedContext) -> None:
        self._accumulator.sync()

    def compute(self) -> torch.Tensor:
        return self._aggregator(self._accumulator.state)

    def reset(self) -> None:
        self._accumulator.reset()

    def to(self, device: str | torch.device | int) -> None:
        self._accumulator.to(device)

    def state_dict(self) -> dict[str, Any]:
        return self._accumulator.state_dict()

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self._accumulator.load_state_dict(state_dict)


class ConfusionMatrixMetricBuilder:
    """Fluent configuration: problem type (``binary / multiclass / multilabel``) -> statistic (``with_*``) ->
    aggregation (``micro / macro / weighted / per_class``) -> ``build()``.  Each axis can be set once.

    Parity: reference ``d9d/metric/impl/classification/confusion_matrix.py:23-334``.
    """

    def __init__(self) -> None:
        self._num_outputs: int | None = None
        self._processor: ClassificationPredictionsProcessor | None = None
        self._statistic: ConfusionMatrixStatistic | None = None
        self._aggregation: ClassificationAggregationMethod | None = None

    def _set_problem(self, processor: ClassificationPredictionsProcessor, num_outputs: int) -> Self:
        if self._processor is not None:
            raise ValueError("A problem type (binary, multiclass, multilabel) has already been configured. "
                             "You cannot chain multiple problem definitions.")
        self._processor, self._num_outputs = processor, num_outputs
        return self

    def binary(self, threshold: float = 0.5) -> Self:
        self._set_problem(ThresholdProcessor(threshold), 1)
        self._aggregation = ClassificationAggregationMethod.MICRO
        return self

    def multiclass(self, num_classes: int, top_k: int | None = None) -> Self:
        if top_k is not None:
            self._set_problem(TopKProcessor(top_k), 1)
            self._aggregation = ClassificationAggregationMethod.MICRO
            return self
        return self._set_problem(OneHotProcessor(num_classes), num_classes)

    def multilabel(self, num_classes: int, threshold: float = 0.5) -> Self:
        return self._set_problem(ThresholdProcessor(threshold), num_classes)

    def with_statistic(self, statistic: ConfusionMatrixStatistic) -> Self:
        if self._statistic is not None:
            raise ValueError("A target statistic has already been configured. "
                             "You cannot evaluate multiple primary statistics in a single pipeline.")
        self._statistic = statistic
        return self

    def with_accuracy(self) -> Self:
        return self.with_statistic(AccuracyStatistic())

    def with_f1(self) -> Self:
        return self.with_statistic(FBetaStatistic(beta=1))

    def with_fbeta(self, beta: float) -> Self:
        return self.with_statistic(FBetaStatistic(beta))

    def with_precision(self) -> Self:
        return self.with_statistic(PrecisionStatistic())

    def with_recall(self) -> Self:
        return self.with_statistic(RecallStatistic())

    def with_aggregation(self, method: ClassificationAggregationMethod) -> Self:
        if self._aggregation is not None:
            raise ValueError("An aggregation methodology has already been selected.")
        self._aggregation = method
        return self

    def micro(self) -> Self:
        return self.with_aggregation(ClassificationAggregationMethod.MICRO)

    def macro(self) -> Self:
        return self.with_aggregation(ClassificationAggregationMethod.MACRO)

    def weighted(self) -> Self:
        return self.with_aggregation(ClassificationAggregationMethod.WEIGHTED)

    def per_class(self) -> Self:
        return self.with_aggregation(ClassificationAggregationMethod.NONE)

    def build(self) -> ConfusionMatrixMetric:
        if self._processor is None or self._num_outputs is None:
            raise ValueError("A problem type (binary, multiclass, multilabel) must be configured.")
        if self._statistic is None:
            raise ValueError("A statistic calculation strategy must be configured.")
        if self._aggregation is None:
            raise ValueError("Aggregation method must be configured.")
        return ConfusionMatrixMetric(self._processor, ConfusionMatrixAccumulator(self._num_outputs),
                                     ConfusionMatrixAggregator(self._aggregation, self._statistic))


def confusion_matrix_metric() -> ConfusionMatrixMetricBuilder:
    return ConfusionMatrixMetricBuilder()
