"""Ready-made providers / tasks for common jobs (used by ``example/`` and ``bench.py``)."""

from .causal_lm import (
    CausalLMTask,
    CausalLMPerplexityTask,
    Qwen3MoEModelProvider,
    Qwen3MoEModelProviderConfig,
    SyntheticDataConfig,
    SyntheticDataProvider,
)

from .throughput import ThroughputMeter, transformer_flops_per_token

__all__ = [
    "CausalLMPerplexityTask",
    "CausalLMTask",
    "Qwen3MoEModelProvider",
    "Qwen3MoEModelProviderConfig",
    "SyntheticDataConfig",
    "SyntheticDataProvider",
    "ThroughputMeter",
    "transformer_flops_per_token",
]
