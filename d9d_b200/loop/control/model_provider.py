from __future__ import annotations

import abc
import dataclasses
from typing import Generic, TypeVar

from torch import nn

from d9d_b200.core.dist_context import DistributedContext
from d9d_b200.core.types import ScalarTree
from d9d_b200.loop.event import EventBus
from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.pipelining.api import PipelineStageInfo

TModel = TypeVar("TModel", bound=nn.Module)


@dataclasses.dataclass(kw_only=True)
class InitializeModelStageContext:
    dist_context: DistributedContext
    stage: PipelineStageInfo


@dataclasses.dataclass(kw_only=True)
class InitializeModelStageResult(Generic[TModel]):
    model: TModel
    state_mapper: ModelStateMapper  # checkpoint keys -> this module's keys


@dataclasses.dataclass(kw_only=True)
class ParallelizeModelStageContext(Generic[TModel]):
    dist_context: DistributedContext
    stage: PipelineStageInfo
    model: TModel


@dataclasses.dataclass(kw_only=True)
class PrepareExportModelStageContext(Generic[TModel]):
    dist_context: DistributedContext
    model: TModel


@dataclasses.dataclass(kw_only=True)
class PrepareExportModelStageResult:
    state_mapper: ModelStateMapper  # this module's keys -> exported keys


@dataclasses.dataclass(kw_only=True)
class RegisterModelEventsContext:
    dist_context: DistributedContext
    event_bus: EventBus


class ModelProvider(abc.ABC, Generic[TModel]):
    """Builds one pipeline stage of the model (on the meta device), parallelises it in place and describes how its
    state maps to / from checkpoints."""

    @abc.abstractmethod
    def initialize_model_stage(self, context: InitializeModelStageContext) -> InitializeModelStageResult[TModel]: ...

    @abc.abstractmethod
    def parallelize_model_stage(self, context: ParallelizeModelStageContext[TModel]) -> None: ...

    @abc.abstractmethod
    def prepare_export_model_stage(self, context: PrepareExportModelStageContext[TModel]) -> PrepareExportModelStageResult: ...

    def register_events(self, context: RegisterModelEventsContext) -> None:  # noqa: B027
        """Optional: subscribe to loop events."""

    def dump_hparams(self) -> ScalarTree:
        return {}
