from typing import Annotated, Literal

from pydantic import BaseModel, Field

from d9d_b200.peft.full_tune.config import FullTuneConfig
from d9d_b200.peft.lora.config import LoRAConfig


class PeftStackConfig(BaseModel):
    kind: Literal["stack"] = "stack"
    methods: list["AnyPeftConfig"]


AnyPeftConfig = Annotated[LoRAConfig | FullTuneConfig | PeftStackConfig, Field(discriminator="kind")]
PeftStackConfig.model_rebuild()
