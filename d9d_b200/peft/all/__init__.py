"""Stack of PEFT methods + config-driven factory."""

from .config import AnyPeftConfig, PeftStackConfig
from .method import PeftStack, peft_method_from_config

__all__ = ["AnyPeftConfig", "PeftStack", "PeftStackConfig", "peft_method_from_config"]
