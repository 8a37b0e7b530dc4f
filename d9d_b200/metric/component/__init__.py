from .accumulator import MetricAccumulator, MetricReduceOp, sync_accumulators

__all__ = ["MetricAccumulator", "MetricReduceOp", "sync_accumulators"]
