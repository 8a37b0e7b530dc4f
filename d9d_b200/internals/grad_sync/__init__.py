"""Bucketed gradient synchronisation keyed on DTensor placements (reference ``d9d/internals/grad_sync``)."""

from .bucket import AbstractGradientBucket, LocalGradientBucket, SyncGradientBucket
from .synchronizer import GradientSynchronizer

__all__ = ["AbstractGradientBucket", "GradientSynchronizer", "LocalGradientBucket", "SyncGradientBucket"]
