from __future__ import annotations

import importlib.util
import sys
from typing import Annotated

from pydantic import Field

from .base import BaseTracker
from .provider.aim.config import AimConfig
from .provider.jsonl import JsonlTracker, JsonlTrackerConfig
from .provider.null import NullTracker, NullTrackerConfig

AnyTrackerConfig = Annotated[AimConfig | NullTrackerConfig | JsonlTrackerConfig, Field(discriminator="provider")]


def tracker_from_config(config: AimConfig | NullTrackerConfig | JsonlTrackerConfig) -> BaseTracker:
    """Instantiate the backend selected by ``config.provider``; optional backends fail with a clear ImportError."""
    if isinstance(config, NullTrackerConfig):
        return NullTracker.from_config(config)
    if isinstance(config, JsonlTrackerConfig):
        return JsonlTracker.from_config(config)
    if isinstance(config, AimConfig):
        if "aim" not in sys.modules and importlib.util.find_spec("aim") is None:  # fail at configuration time, not at the first step
            raise ImportError(f"The tracker configuration {config.provider} could not be loaded - ensure these "
                              "dependencies are installed: aim")
        from .provider.aim.tracker import AimTracker

        return AimTracker.from_config(config)
    raise TypeError(f"unknown tracker config {type(config).__name__}")
