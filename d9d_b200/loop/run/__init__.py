"""Entry points: ``TrainingConfigurator(...).configure() -> Trainer`` and ``InferenceConfigurator(...).configure() -> Inference``."""

from .inference import Inference, InferenceConfigurator
from .train import Trainer, TrainingConfigurator

__all__ = ["Inference", "InferenceConfigurator", "Trainer", "TrainingConfigurator"]
