from re import Pattern
from typing import Literal

from pydantic import BaseModel


class FullTuneConfig(BaseModel):
    kind: Literal["full_tune"] = "full_tune"
    module_name_pattern: Pattern
