"""Pydantic job configuration (reference ``d9d/loop/config``)."""

from .config import (
    BatchingConfig,
    CheckpointingConfig,
    DataLoadingConfig,
    DeterminismConfig,
    GarbageCollectionConfig,
    GradientClippingConfig,
    GradientManagerConfig,
    InferenceConfig,
    JobLoggerConfig,
    Fp8LinearConfig,
    ModelStageFactoryConfig,
    PipeliningConfig,
    ProfilingConfig,
    TimeoutConfig,
    TrainerConfig,
)
from .types import StepActionPeriod, StepActionSpecial

__all__ = [
    "BatchingConfig",
    "CheckpointingConfig",
    "DataLoadingConfig",
    "DeterminismConfig",
    "GarbageCollectionConfig",
    "GradientClippingConfig",
    "GradientManagerConfig",
    "InferenceConfig",
    "JobLoggerConfig",
    "Fp8LinearConfig",
    "ModelStageFactoryConfig",
    "PipeliningConfig",
    "ProfilingConfig",
    "StepActionPeriod",
    "StepActionSpecial",
    "TimeoutConfig",
    "TrainerConfig",
]
