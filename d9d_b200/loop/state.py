"""Aggregated job state (everything the loops need + what gets checkpointed).

Checkpoint key tree (reference ``d9d/loop/state.py:64-118``): ``stepper``, ``tracked_modules``, ``data_loader`` and,
for training, ``logger``, ``task``, ``metrics``, ``optimizer``, ``lr_scheduler``.
"""

from __future__ import annotations

import dataclasses
from typing import Any

from torch.distributed.checkpoint.stateful import Stateful
from torchdata.stateful_dataloader import StatefulDataLoader

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.core.protocol import LRSchedulerProtocol, OptimizerProtocol
from d9d_b200.loop.component import (
    BatchMaths,
    GradientClipper,
    GradientManager,
    InferenceTaskOperator,
    JobLogger,
    JobProfiler,
    ManualGarbageCollector,
    ModelStageExporter,
    StateCheckpointer,
    Stepper,
    TimeoutManager,
    TrackedModules,
    TrainTaskOperator,
)
from d9d_b200.loop.control import InferenceTask, TrainTask
from d9d_b200.loop.event import EventBus
from d9d_b200.metric.impl.container import ComposeMetric


@dataclasses.dataclass(kw_only=True)
class JobState(Stateful):
    dist_context: DistributedContext
    stepper: Stepper
    garbage_collector: ManualGarbageCollector
    checkpointer: StateCheckpointer
    profiler: JobProfiler
    tracked_modules: TrackedModules
    batch_maths: BatchMaths
    data_loader: StatefulDataLoader
    timeout_manager: TimeoutManager

    def state_dict(self) -> dict[str, Any]:
        return {
            "stepper": self.stepper.state_dict(),
            "tracked_modules": self.tracked_modules.state_dict(),
            "data_loader": self.data_loader.state_dict(),
        }

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        self.stepper.load_state_dict(state_dict["stepper"])
        self.tracked_modules.load_state_dict(state_dict["tracked_modules"])
        self.data_loader.load_state_dict(state_dict["data_loader"])


@dataclasses.dataclass(kw_only=True)
class TrainJobState(JobState):
    task: TrainTask
    gradient_manager: GradientManager
    metrics: ComposeMetric
    task_operator: TrainTaskOperator
    logger: JobLogger
    optimizer: OptimizerProtocol
    lr_scheduler: LRSchedulerProtocol
    gradient_clipper: GradientClipper
    exporter: ModelStageExporter
    event_bus: EventBus

    def state_dict(self) -> dict[str, Any]:
        return {
            **super().state_dict(),
            "logger": self.logger.state_dict(),
            "task": self.task.state_dict(),
            "metrics": self.metrics.state_dict(),
            "optimizer": self.optimizer.state_dict(),
            "lr_scheduler": self.lr_scheduler.state_dict(),
        }

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        super().load_state_dict(state_dict)
        self.logger.load_state_dict(state_dict["logger"])
        self.task.load_state_dict(state_dict["task"])
        self.metrics.load_state_dict(state_dict["metrics"])
        self.optimizer.load_state_dict(state_dict["optimizer"])
        self.lr_scheduler.load_state_dict(state_dict["lr_scheduler"])


@dataclasses.dataclass(kw_only=True)
class InferenceJobState(JobState):
    task: InferenceTask
    task_operator: InferenceTaskOperator
    event_bus: EventBus

    def state_dict(self) -> dict[str, Any]:
        return {**super().state_dict(), "task": self.task.state_dict()}

    def load_state_dict(self, state_dict: dict[str, Any]) -> None:
        super().load_state_dict(state_dict)
        self.task.load_state_dict(state_dict["task"])
