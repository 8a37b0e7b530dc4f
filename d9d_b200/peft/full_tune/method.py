from __future__ import annotations

from typing import Self

from torch import nn

from ..base import PeftInjectionResult, PeftMethod
from .config import FullTuneConfig


class FullTune(PeftMethod[FullTuneConfig]):
    """No structural change: the parameters of every module whose qualified name full-matches the pattern become the
    trainable set (each parameter once, even if several matching modules share it)."""

    def __init__(self, config: FullTuneConfig):
        self._pattern = config.module_name_pattern

    @classmethod
    def from_config(cls, config: FullTuneConfig) -> Self:
        return cls(config)

    def inject(self, module: nn.Module) -> PeftInjectionResult:
        selected: dict[int, nn.Parameter] = {}
        for name, sub in module.named_modules():
            if self._pattern.fullmatch(name) is None:
                continue
            for p in sub.parameters():
                selected.setdefault(id(p), p)
        return PeftInjectionResult(parameters_to_train=list(selected.values()), load_state_mappers=[])

    def merge(self, module: nn.Module) -> None:
        """Nothing was injected, so there is nothing to fold back."""
