"""High-level helpers applying a parallelism strategy to a module (reference ``d9d/module/parallelism/api``)."""

from .context_parallel import parallelize_context_parallel
from .expert_parallel import parallelize_expert_parallel
from .fully_sharded import parallelize_fsdp
from .hybrid_sharded import parallelize_hsdp
from .replicate_parallel import parallelize_replicate
from .tensor_parallel import (
    parallelize_attention_tensor_parallel,
    parallelize_colwise,
    parallelize_rowwise,
    parallelize_swiglu_tensor_parallel,
)

__all__ = [
    "parallelize_attention_tensor_parallel",
    "parallelize_colwise",
    "parallelize_context_parallel",
    "parallelize_expert_parallel",
    "parallelize_fsdp",
    "parallelize_hsdp",
    "parallelize_replicate",
    "parallelize_rowwise",
    "parallelize_swiglu_tensor_parallel",
]
