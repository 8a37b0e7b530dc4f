"""Low-rank adaptation for ``nn.Linear`` and ``GroupedLinear`` layers."""

from .config import LoRAConfig, LoRAParameters
from .layer import LoRAGroupedLinear, LoRALinear
from .method import LoRA

__all__ = ["LoRA", "LoRAConfig", "LoRAGroupedLinear", "LoRALinear", "LoRAParameters"]
