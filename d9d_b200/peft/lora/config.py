from re import Pattern
from typing import Literal

from pydantic import BaseModel


class LoRAParameters(BaseModel):
    r: int
    alpha: int
    dropout: float


class LoRAConfig(BaseModel):
    kind: Literal["lora"] = "lora"
    module_name_pattern: Pattern  # full-matched against module names
    params: LoRAParameters
