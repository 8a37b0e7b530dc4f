from __future__ import annotations

from d9d_b200.core.dist_context import BATCH_DOMAIN, DistributedContext
from d9d_b200.loop.config import BatchingConfig, PipeliningConfig


class BatchMaths:
    """Derives per-rank batch quantities from ``global_batch_size`` / ``microbatch_size`` and the mesh.

    ``n = gbs / (dp * mbs)`` microbatches per optimizer step per replica: with pipeline parallelism they are the
    schedule's microbatches (the loader yields one batch of ``mbs * n``), without it they are gradient-accumulation
    rounds (the loader yields groups of ``n`` batches of ``mbs``).  Parity: reference ``batch_maths.py:5-109``.
    """

    def __init__(self, dist_context: DistributedContext, config_batching: BatchingConfig, config_pipelining: PipeliningConfig | None):
        self._ctx = dist_context
        self._batching = config_batching
        self._pipelining = config_pipelining
        dp = dist_context.mesh_for(BATCH_DOMAIN)["dp"].size() if dist_context.mesh_params.is_distributed else 1
        self._dp_size = dp
        self._global_microbatch = dp * config_batching.microbatch_size
        if config_batching.global_batch_size % self._global_microbatch != 0:
            raise ValueError("Global Batch Size must be divisible by (Data Parallel cardinality * Microbatch Size)")

    @property
    def global_batch_size(self) -> int:
        return self._batching.global_batch_size

    @property
    def microbatch_size(self) -> int:
        return self._batching.microbatch_size

    @property
    def data_parallel_size(self) -> int:
        return self._dp_size

    @property
    def _microbatches_per_step(self) -> int:
        return self._batching.global_batch_size // self._global_microbatch

    @property
    def num_microbatches_pipelining(self) -> int:
        return self._microbatches_per_step if self._ctx.mesh_params.has_pipeline_parallel else 1

    @property
    def num_microbatches_gradient_accumulation(self) -> int:
        return 1 if self._ctx.mesh_params.has_pipeline_parallel else self._microbatches_per_step

    @property
    def data_loader_batch_size(self) -> int:
        return self._batching.microbatch_size * self.num_microbatches_pipelining

    @property
    def num_backward_calls(self) -> int:
        return self.num_microbatches_pipelining * self.num_microbatches_gradient_accumulation
