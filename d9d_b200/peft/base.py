from __future__ import annotations

import abc
import dataclasses
from typing import Generic, Self, TypeVar

from pydantic import BaseModel
from torch import nn

from d9d_b200.model_state.mapper import ModelStateMapper


@dataclasses.dataclass(slots=True)
class PeftInjectionResult:
    parameters_to_train: list[nn.Parameter]
    #: mappers that let an *unmodified* checkpoint load into the modified module tree
    load_state_mappers: list[ModelStateMapper]


TConfig = TypeVar("TConfig", bound=BaseModel)


class PeftMethod(abc.ABC, Generic[TConfig]):
    @abc.abstractmethod
    def inject(self, module: nn.Module) -> PeftInjectionResult: ...

    @abc.abstractmethod
    def merge(self, module: nn.Module) -> None: ...

    @classmethod
    @abc.abstractmethod
    def from_config(cls, config: TConfig) -> Self: ...
