"""Experiment tracking: backend-agnostic run interface + providers (reference ``d9d/tracker``)."""

from .base import BaseTracker, BaseTrackerRun, RunConfig
from .factory import AnyTrackerConfig, tracker_from_config

__all__ = ["AnyTrackerConfig", "BaseTracker", "BaseTrackerRun", "RunConfig", "tracker_from_config"]
