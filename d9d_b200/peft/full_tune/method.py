from __future__ import annotations

from typing import Self

from torch import nn

from ..base import PeftInjectionResult, PeftMethod
from .config import FullTuneConfig


class FullTune(PeftMethod[FullTuneConfig]):
    """Marks every parameter of the modules whose name full-matches the pattern as trainable; no structural change."""

    def __init__(self, config: FullTuneConfig):
        self._config = config

    def inject(self, module: nn.Module) -> PeftInjectionResult:
        train: list[nn.Parameter] = []
        for name, mod in module.named_modules():
            if self._config.module_name_pattern.fullmatch(name):
                train.extend(mod.parameters())
        return PeftInjectionResult(parameters_to_train=train, load_state_mappers=[])

    def merge(self, module: nn.Module) -> None:
        return None

    @classmethod
    def from_config(cls, config: FullTuneConfig) -> Self:
        return cls(config)
