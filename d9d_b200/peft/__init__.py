"""Parameter-efficient fine-tuning: inject adapters, freeze the rest, merge back (reference ``d9d/peft``)."""

from .applicator import inject_peft_and_freeze, merge_peft
from .base import PeftInjectionResult, PeftMethod

__all__ = ["PeftInjectionResult", "PeftMethod", "inject_peft_and_freeze", "merge_peft"]
