"""Asynchronous metric synchronisation/computation on a side stream (reference ``internals/metric_collector``)."""

from .collector import AsyncMetricCollector

__all__ = ["AsyncMetricCollector"]
