"""User extension points: model / dataset / optimizer / LR-scheduler providers and tasks
(reference ``d9d/loop/control``)."""

from .dataset_provider import DatasetProvider, InitializeDatasetContext, InitializeDatasetResult
from .lr_scheduler_provider import InitializeLRSchedulerContext, LRSchedulerProvider
from .model_provider import (
    InitializeModelStageContext,
    InitializeModelStageResult,
    ModelProvider,
    ParallelizeModelStageContext,
    PrepareExportModelStageContext,
    PrepareExportModelStageResult,
    RegisterModelEventsContext,
)
from .optimizer_provider import InitializeOptimizerStageContext, OptimizerProvider
from .task import (
    BaseTask,
    BuildForwardInputsContext,
    BuildForwardInputsResult,
    ComputeLossContext,
    ComputeLossResult,
    CreateMetricsContext,
    CreateMetricsResult,
    FinalizeContext,
    InferenceTask,
    InferenceTaskProvider,
    InferenceTaskProviderContext,
    ProcessOutputsContext,
    RegisterTaskEventsContext,
    TrainTask,
    TrainTaskProvider,
    TrainTaskProviderContext,
    UpdateMetricsContext,
)

__all__ = [
    "BaseTask",
    "BuildForwardInputsContext",
    "BuildForwardInputsResult",
    "ComputeLossContext",
    "ComputeLossResult",
    "CreateMetricsContext",
    "CreateMetricsResult",
    "DatasetProvider",
    "FinalizeContext",
    "InferenceTask",
    "InferenceTaskProvider",
    "InferenceTaskProviderContext",
    "InitializeDatasetContext",
    "InitializeDatasetResult",
    "InitializeLRSchedulerContext",
    "InitializeModelStageContext",
    "InitializeModelStageResult",
    "InitializeOptimizerStageContext",
    "LRSchedulerProvider",
    "ModelProvider",
    "OptimizerProvider",
    "ParallelizeModelStageContext",
    "PrepareExportModelStageContext",
    "PrepareExportModelStageResult",
    "ProcessOutputsContext",
    "RegisterModelEventsContext",
    "RegisterTaskEventsContext",
    "TrainTask",
    "TrainTaskProvider",
    "TrainTaskProviderContext",
    "UpdateMetricsContext",
]
