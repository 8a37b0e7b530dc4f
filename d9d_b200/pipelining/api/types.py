from collections.abc import Callable
from typing import Any

import torch

PipelineResultFn = Callable[[dict[str, torch.Tensor], int], Any]
"""``(last stage outputs, microbatch index) -> anything`` (forward-only schedules)."""

PipelineLossFn = Callable[[dict[str, torch.Tensor], int], torch.Tensor]
"""``(last stage outputs, microbatch index) -> scalar loss`` to back-propagate."""
