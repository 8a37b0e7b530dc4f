"""Distributed metrics (reference ``d9d/metric``)."""

from .abc import Metric

__all__ = ["Metric"]
