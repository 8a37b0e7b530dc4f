"""Structural protocols for the training objects the loop manipulates (reference ``d9d/core/protocol/training.py``)."""

from .training import LRSchedulerProtocol, OptimizerProtocol

__all__ = ["LRSchedulerProtocol", "OptimizerProtocol"]
