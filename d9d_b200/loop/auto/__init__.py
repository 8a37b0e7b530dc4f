"""Config-driven optimizer / LR-scheduler providers (reference ``d9d/loop/auto``)."""

from .auto_lr_scheduler import AutoLRSchedulerConfig, AutoLRSchedulerProvider
from .auto_optimizer import AutoOptimizerConfig, AutoOptimizerProvider

__all__ = ["AutoLRSchedulerConfig", "AutoLRSchedulerProvider", "AutoOptimizerConfig", "AutoOptimizerProvider"]
