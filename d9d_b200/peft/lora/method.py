from __future__ import annotations

from typing import Self

import torch
from torch import nn

from d9d_b200.model_state.mapper import ModelStateMapper
from d9d_b200.model_state.mapper.leaf import ModelStateMapperRename
from d9d_b200.module.block.moe import GroupedLinear

from ..base import PeftInjectionResult, PeftMethod
from .config import LoRAConfig
from .layer import LoRAGroupedLinear, LoRALinear

_WRAPPERS = (LoRALinear, LoRAGroupedLinear)


def _candidates(root: nn.Module) -> list[tuple[str, nn.Module]]:
    """(name, module) of every Linear / GroupedLinear that is not already inside a LoRA wrapper."""
    found: list[tuple[str, nn.Module]] = []
    stack: list[tuple[str, nn.Module]] = [("", root)]
    seen: set[int] = set()
    while stack:
        name, mod = stack.pop()
        if id(mod) in seen or isinstance(mod, _WRAPPERS):
            continue
        seen.add(id(mod))
        if isinstance(mod, (nn.Linear, GroupedLinear)):
            found.append((name, mod))
        for child_name, child in reversed(list(mod.named_children())):
            stack.append((f"{name}.{child_name}" if name else child_name, child))
    return found


class LoRA(PeftMethod[LoRAConfig]):
    """Wrap every matching ``nn.Linear`` / ``GroupedLinear`` and emit ``x.weight -> x.base.weight`` renames so stock
    checkpoints keep loading (reference ``peft/lora/method.py:18-125``)."""

    def __init__(self, config: LoRAConfig):
        self._config = config

    def inject(self, module: nn.Module) -> PeftInjectionResult:
        train: list[nn.Parameter] = []
        mappers: list[ModelStateMapper] = []
        for name, mod in _candidates(module):
            if not self._config.module_name_pattern.fullmatch(name):
                continue
            wrapper = LoRALinear(mod, self._config.params) if isinstance(mod, nn.Linear) else LoRAGroupedLinear(mod, self._config.params)
            train += list(wrapper.lora_A.parameters()) + list(wrapper.lora_B.parameters())
            mappers.append(ModelStateMapperRename(name_from=f"{name}.weight", name_to=f"{name}.base.weight"))
            module.set_submodule(name, wrapper)
        return PeftInjectionResult(parameters_to_train=train, load_state_mappers=mappers)

    def merge(self, module: nn.Module) -> None:
        for name, mod in list(module.named_modules()):
            if isinstance(mod, _WRAPPERS) and self._config.module_name_pattern.fullmatch(name):
                with torch.no_grad():
                    module.set_submodule(name, mod.merge_with_base_())

    @classmethod
    def from_config(cls, config: LoRAConfig) -> Self:
        return cls(config)
