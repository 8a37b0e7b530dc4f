from __future__ import annotations

from typing import Any, Self

from torch import nn

from ..base import PeftInjectionResult, PeftMethod
from ..full_tune.config import FullTuneConfig
from ..full_tune.method import FullTune
from ..lora.config import LoRAConfig
from ..lora.method import LoRA
from .config import PeftStackConfig


class PeftStack(PeftMethod[PeftStackConfig]):
    """Applies methods in order; merges in reverse order."""

    def __init__(self, methods: list[PeftMethod]):
        self._methods = methods

    def inject(self, module: nn.Module) -> PeftInjectionResult:
        train, mappers = [], []
        for m in self._methods:
            res = m.inject(module)
            train += res.parameters_to_train
            mappers += res.load_state_mappers
        return PeftInjectionResult(parameters_to_train=train, load_state_mappers=mappers)

    def merge(self, module: nn.Module) -> None:
        for m in reversed(self._methods):
            m.merge(module)

    @classmethod
    def from_config(cls, config: PeftStackConfig) -> Self:
        return cls([peft_method_from_config(c) for c in config.methods])


def peft_method_from_config(config: Any) -> PeftMethod:
    if isinstance(config, LoRAConfig):
        return LoRA.from_config(config)
    if isinstance(config, FullTuneConfig):
        return FullTune.from_config(config)
    if isinstance(config, PeftStackConfig):
        return PeftStack.from_config(config)
    raise TypeError(f"unknown PEFT config {type(config).__name__}")
